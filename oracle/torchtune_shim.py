"""ORACLE — test infrastructure only (never imported by the product path).

CPU restatement of the arithmetic of ``torchtune==0.4.0`` (pinned at
/root/reference/pyproject.toml:19) that the reference's ``Model`` is built on
(/root/reference/src/csm/models/model.py:7-8,13-25,30-42).  torchtune is a
third-party dependency that is NOT vendored under /root/reference and is not
installable here (no network), so its published algorithm is restated:

  * RMSNorm            : fp32 normalise -> cast back -> multiply by ``scale``
  * Llama3ScaledRoPE   : Llama-3.1 frequency scaling (low=1, high=4, old ctx 8192),
                         interleaved (adjacent-pair) rotation computed in fp32
  * MultiHeadAttention : GQA with each KV head shared by H/KV *adjacent* q heads,
                         F.scaled_dot_product_attention, scale 1/sqrt(hd)
  * FeedForward        : w2(silu(w1 x) * w3 x)
  * layer              : h = attn(sa_norm(x)) + x ; out = h + mlp(mlp_norm(h))
  * TransformerDecoder : tok_embeddings -> layers -> norm -> output(h).float()

Parity of this restatement is UNPINNED by the reference (it ships no golden
vectors for this path).  It is cross-checked in
tests/test_oracle.py against transformers' independent Llama-3 RoPE scaling and
against a dense-matrix attention restatement.

``install()`` injects the shim as ``torchtune`` into ``sys.modules`` so that the
reference's own ``model.py`` can be imported *verbatim* from /root/reference when
it is present (this container only) — see oracle/reference_loader.py.
"""
from __future__ import annotations

import math
import sys
import types
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.scale = nn.Parameter(torch.ones(dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x32 = x.float()
        xn = (x32 * torch.rsqrt(x32.pow(2).mean(-1, keepdim=True) + self.eps)).type_as(x)
        return xn * self.scale


def llama3_scaled_freqs(head_dim: int, base: float, scale_factor: float,
                        low_freq_factor: float = 1.0, high_freq_factor: float = 4.0,
                        old_context_len: int = 8192) -> torch.Tensor:
    """theta_j for j < head_dim/2, after Llama-3 frequency scaling (fp32)."""
    freqs = 1.0 / (base ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    low_wavelen = old_context_len / low_freq_factor
    high_wavelen = old_context_len / high_freq_factor
    out = []
    for f in freqs.tolist():
        wavelen = 2 * math.pi / f
        if wavelen < high_wavelen:
            out.append(f)
        elif wavelen > low_wavelen:
            out.append(f / scale_factor)
        else:
            smooth = (old_context_len / wavelen - low_freq_factor) / (high_freq_factor - low_freq_factor)
            out.append((1 - smooth) * f / scale_factor + smooth * f)
    return torch.tensor(out, dtype=freqs.dtype)


class Llama3ScaledRoPE(nn.Module):
    def __init__(self, dim: int, max_seq_len: int = 4096, base: float = 500_000.0,
                 scale_factor: float = 8.0):
        super().__init__()
        self.dim, self.base, self.max_seq_len, self.scale_factor = dim, base, max_seq_len, scale_factor
        theta = llama3_scaled_freqs(dim, base, scale_factor)
        seq_idx = torch.arange(max_seq_len, dtype=theta.dtype)
        idx_theta = torch.einsum("i,j->ij", seq_idx, theta).float()
        cache = torch.stack([torch.cos(idx_theta), torch.sin(idx_theta)], dim=-1)
        self.register_buffer("cache", cache, persistent=False)  # [max_seq, hd/2, 2] fp32

    def forward(self, x: torch.Tensor, *, input_pos: Optional[torch.Tensor] = None) -> torch.Tensor:
        # x: [b, s, n_h, h_d]
        seq_len = x.size(1)
        rope_cache = self.cache[:seq_len] if input_pos is None else self.cache[input_pos]
        xs = x.float().reshape(*x.shape[:-1], -1, 2)
        rope_cache = rope_cache.view(-1, xs.size(1), 1, xs.size(3), 2)
        out = torch.stack(
            [xs[..., 0] * rope_cache[..., 0] - xs[..., 1] * rope_cache[..., 1],
             xs[..., 1] * rope_cache[..., 0] + xs[..., 0] * rope_cache[..., 1]], -1)
        return out.flatten(3).type_as(x)


class MultiHeadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, num_kv_heads, head_dim, pos_embeddings, max_seq_len):
        super().__init__()
        self.num_heads, self.num_kv_heads, self.head_dim = num_heads, num_kv_heads, head_dim
        self.embed_dim, self.max_seq_len = embed_dim, max_seq_len
        self.q_proj = nn.Linear(embed_dim, num_heads * head_dim, bias=False)
        self.k_proj = nn.Linear(embed_dim, num_kv_heads * head_dim, bias=False)
        self.v_proj = nn.Linear(embed_dim, num_kv_heads * head_dim, bias=False)
        self.output_proj = nn.Linear(embed_dim, embed_dim, bias=False)
        self.pos_embeddings = pos_embeddings
        self.kv_cache = None        # inference only: (k [b, H, max_seq, hd], v, filled) — torchtune 0.4.0 KVCache

    def forward(self, x, y, *, mask=None, input_pos=None):
        b, s, _ = x.shape
        rep = self.num_heads // self.num_kv_heads
        q = self.q_proj(x).view(b, s, self.num_kv_heads * rep, self.head_dim)
        q = self.pos_embeddings(q, input_pos=input_pos).transpose(1, 2)
        k = self.k_proj(y).view(b, s, -1, self.head_dim)
        v = self.v_proj(y).view(b, s, -1, self.head_dim)
        k = self.pos_embeddings(k, input_pos=input_pos)
        k = k.view(b, s, self.num_kv_heads, 1, self.head_dim).expand(b, s, self.num_kv_heads, rep, self.head_dim)
        v = v.view(b, s, self.num_kv_heads, 1, self.head_dim).expand(b, s, self.num_kv_heads, rep, self.head_dim)
        k = k.reshape(b, s, -1, self.head_dim).transpose(1, 2)
        v = v.reshape(b, s, -1, self.head_dim).transpose(1, 2)
        if self.kv_cache is not None:
            # torchtune 0.4.0 KVCache.update: the (head-expanded, rotated) keys / values of this call are appended
            # behind what the cache already holds; attention then runs over the WHOLE cache with the caller's
            # [b, s, max_seq] mask (model.py:165,183: rows of the tril mask picked by input_pos)
            kc, vc, filled = self.kv_cache
            if filled + s > kc.shape[2]:
                raise ValueError("kv cache overflow")
            kc[:b, :, filled:filled + s] = k
            vc[:b, :, filled:filled + s] = v
            self.kv_cache = (kc, vc, filled + s)
            k, v = kc[:b], vc[:b]
        if mask is not None:
            mask = mask[:, None, :, :]
        out = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0,
                                             is_causal=(self.kv_cache is None and mask is None))
        out = out.transpose(1, 2).contiguous().view(b, s, -1)
        return self.output_proj(out)


class FeedForward(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.w1 = nn.Linear(dim, hidden, bias=False)   # gate
        self.w2 = nn.Linear(hidden, dim, bias=False)   # down
        self.w3 = nn.Linear(dim, hidden, bias=False)   # up

    def forward(self, x):
        return self.w2(F.silu(self.w1(x)) * self.w3(x))


class TransformerSelfAttentionLayer(nn.Module):
    def __init__(self, attn, mlp, sa_norm, mlp_norm):
        super().__init__()
        self.attn, self.mlp, self.sa_norm, self.mlp_norm = attn, mlp, sa_norm, mlp_norm

    def forward(self, x, *, mask=None, input_pos=None):
        n = self.sa_norm(x)
        h = self.attn(n, n, mask=mask, input_pos=input_pos) + x
        return h + self.mlp(self.mlp_norm(h))


class TransformerDecoder(nn.Module):
    def __init__(self, *, tok_embeddings, layers, max_seq_len, num_heads, head_dim, norm, output):
        super().__init__()
        self.tok_embeddings = tok_embeddings
        self.layers = nn.ModuleList(layers)
        self.norm, self.output = norm, output
        self.max_seq_len, self.num_heads, self.head_dim = max_seq_len, num_heads, head_dim
        self._caches = False

    # KV caches are an inference feature (generate_frame, model.py:140-195); the training oracle never enables them.
    def setup_caches(self, batch_size, dtype, *, encoder_max_seq_len=None, decoder_max_seq_len=None):
        n = decoder_max_seq_len if decoder_max_seq_len is not None else self.max_seq_len
        for layer in self.layers:
            a = layer.attn
            dev = a.q_proj.weight.device
            a.kv_cache = (torch.zeros(batch_size, a.num_heads, n, a.head_dim, dtype=dtype, device=dev),
                          torch.zeros(batch_size, a.num_heads, n, a.head_dim, dtype=dtype, device=dev), 0)
        self._caches = True

    def caches_are_enabled(self):
        return self._caches

    def reset_caches(self):
        if not self._caches:
            raise RuntimeError("Key value caches are not setup. Call ``setup_caches()`` first.")
        for layer in self.layers:
            kc, vc, _ = layer.attn.kv_cache
            layer.attn.kv_cache = (kc.zero_(), vc.zero_(), 0)

    def forward(self, tokens, *, mask=None, input_pos=None):
        seq_len = tokens.shape[1]
        if seq_len > self.max_seq_len:
            raise ValueError(f"seq_len ({seq_len}) of input tensor should be smaller "
                             f"than max_seq_len ({self.max_seq_len})")
        h = self.tok_embeddings(tokens)
        for layer in self.layers:
            h = layer(h, mask=mask, input_pos=input_pos)
        h = self.norm(h)
        return self.output(h).float()


def llama3_2(vocab_size, num_layers, num_heads, num_kv_heads, embed_dim, max_seq_len=131072,
             attn_dropout=0.0, rope_base=500_000, intermediate_dim=None, norm_eps=1e-5,
             scale_factor=32) -> TransformerDecoder:
    head_dim = embed_dim // num_heads
    num_kv_heads = num_kv_heads if num_kv_heads else num_heads
    rope = Llama3ScaledRoPE(dim=head_dim, max_seq_len=max_seq_len, base=rope_base, scale_factor=scale_factor)
    layers = []
    for _ in range(num_layers):
        attn = MultiHeadAttention(embed_dim, num_heads, num_kv_heads, head_dim, rope, max_seq_len)
        mlp = FeedForward(embed_dim, intermediate_dim)
        layers.append(TransformerSelfAttentionLayer(attn, mlp, RMSNorm(embed_dim, norm_eps),
                                                    RMSNorm(embed_dim, norm_eps)))
    # The real builder allocates nn.Embedding(vocab_size, embed_dim) and a tied output
    # projection; the reference replaces both with nn.Identity (model.py:51-56) and reads
    # only ``tok_embeddings.embedding_dim``.  A meta-device embedding keeps that attribute
    # without allocating 128256 x 2048 floats that are discarded immediately.
    tok = nn.Embedding(vocab_size, embed_dim, device="meta")
    return TransformerDecoder(tok_embeddings=tok, layers=layers, max_seq_len=max_seq_len,
                              num_heads=num_heads, head_dim=head_dim,
                              norm=RMSNorm(embed_dim, norm_eps), output=nn.Identity())


def install() -> None:
    """Expose this module as ``torchtune`` (only if the real one is absent)."""
    if "torchtune" in sys.modules:
        return
    tt = types.ModuleType("torchtune")
    models = types.ModuleType("torchtune.models")
    l32 = types.ModuleType("torchtune.models.llama3_2")
    modules = types.ModuleType("torchtune.modules")
    transformer = types.ModuleType("torchtune.modules.transformer")
    l32.llama3_2 = llama3_2
    transformer.TransformerDecoder = TransformerDecoder
    modules.transformer = transformer
    modules.RMSNorm = RMSNorm
    models.llama3_2 = l32
    tt.models, tt.modules = models, modules
    sys.modules.update({"torchtune": tt, "torchtune.models": models, "torchtune.models.llama3_2": l32,
                        "torchtune.modules": modules, "torchtune.modules.transformer": transformer})
