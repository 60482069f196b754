"""CSM ``Model`` on libcsm_b200 — drop-in for /root/reference/src/csm/models/model.py.

Same constructor (``ModelArgs``), attributes and state-dict keys as the reference (model.py:99-126;
backbone./decoder. keys follow torchtune 0.4.0: ``layers.{i}.attn.{q,k,v}_proj.weight``,
``layers.{i}.attn.output_proj.weight``, ``layers.{i}.mlp.{w1,w2,w3}.weight``,
``layers.{i}.{sa_norm,mlp_norm}.scale``, ``norm.scale``), same helpers (``_embed_tokens``, ``_embed_audio``,
``_index_causal_mask``, ``_create_causal_mask``).  New: ``Model.forward`` — the training forward the reference only
has as a free function (training/utils.py:56-119) plus the decoder term it leaves as a placeholder
(utils.py:109-117; restated from generate_frame model.py:171-193).

Every arithmetic op runs in a CUDA kernel of libcsm_b200.so (bf16 storage, fp32 accumulation).  There is no CPU
path: calling forward on CPU tensors raises.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from .. import autograd as _ag
from ..autograd import (DecoderInputFn, EmbedGatherSumFn, GroupedLinearCEFn, LinearCEFn, LinearFn, StackFn)
from .rope import build_rope_cache

BF16 = torch.bfloat16


# ----------------------------------------------------------------------------- transformer modules (parameters only)
class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float):
        super().__init__()
        self.eps = eps
        self.scale = nn.Parameter(torch.ones(dim))


class MultiHeadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, num_kv_heads, head_dim):
        super().__init__()
        self.q_proj = nn.Linear(embed_dim, num_heads * head_dim, bias=False)
        self.k_proj = nn.Linear(embed_dim, num_kv_heads * head_dim, bias=False)
        self.v_proj = nn.Linear(embed_dim, num_kv_heads * head_dim, bias=False)
        self.output_proj = nn.Linear(embed_dim, embed_dim, bias=False)


class FeedForward(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.w1 = nn.Linear(dim, hidden, bias=False)   # gate
        self.w2 = nn.Linear(hidden, dim, bias=False)   # down
        self.w3 = nn.Linear(dim, hidden, bias=False)   # up


class TransformerLayer(nn.Module):
    def __init__(self, embed_dim, num_heads, num_kv_heads, head_dim, intermediate_dim, eps):
        super().__init__()
        self.attn = MultiHeadAttention(embed_dim, num_heads, num_kv_heads, head_dim)
        self.mlp = FeedForward(embed_dim, intermediate_dim)
        self.sa_norm = RMSNorm(embed_dim, eps)
        self.mlp_norm = RMSNorm(embed_dim, eps)


class TransformerDecoder(nn.Module):
    """Parameter container with torchtune's TransformerDecoder interface as the reference uses it
    (model.py:53-55,137,169,184): ``tok_embeddings``/``output`` attributes, ``max_seq_len``,
    ``forward(h, input_pos=, mask=) -> fp32``.  Causal attention only (training)."""

    def __init__(self, num_layers, num_heads, num_kv_heads, embed_dim, max_seq_len, intermediate_dim,
                 norm_eps=1e-5, rope_base=500_000.0, scale_factor=32.0):
        super().__init__()
        self.num_heads, self.num_kv_heads = num_heads, num_kv_heads
        self.embed_dim, self.head_dim = embed_dim, embed_dim // num_heads
        self.max_seq_len, self.norm_eps = max_seq_len, norm_eps
        self.rope_base, self.scale_factor = rope_base, scale_factor
        self.tok_embeddings = nn.Identity()
        self.output = nn.Identity()
        self.layers = nn.ModuleList([
            TransformerLayer(embed_dim, num_heads, num_kv_heads, self.head_dim, intermediate_dim, norm_eps)
            for _ in range(num_layers)])
        self.norm = RMSNorm(embed_dim, norm_eps)
        self._rope = {}
        self._packed = {}          # fused qkv / gate-up weight storage (csm/autograd.py::_packed_weight)

    def rope_cache(self, device) -> torch.Tensor:
        key = (str(device), self.max_seq_len)
        if key not in self._rope:
            self._rope[key] = build_rope_cache(self.head_dim, self.max_seq_len, self.rope_base,
                                               self.scale_factor).to(device)
        return self._rope[key]

    def set_max_seq_len(self, n: int) -> None:
        """Llama-3 scaling is position independent, so a longer context only extends the table."""
        self.max_seq_len = n

    # ---- KV caches: inference only (generate_frame, model.py:140-195).  The training path (``run``) never touches
    # them, so the reference trainer's habit of calling setup_caches before training (trainer.py:114,121) is harmless.
    def caches_are_enabled(self) -> bool:
        return getattr(self, "_kc", None) is not None

    def setup_caches(self, max_batch_size: int, dtype=BF16, *, encoder_max_seq_len=None, decoder_max_seq_len=None):
        """torchtune TransformerDecoder.setup_caches as model.py:134-135 calls it: per layer K / V caches
        [max_batch, max_seq, KV*hd] (keys stored rotated; unlike torchtune 0.4.0 the KV heads are NOT expanded to the
        query heads — the decode kernel shares them)."""
        if dtype != BF16:
            raise RuntimeError("csm_b200: KV caches are bf16")
        n = decoder_max_seq_len if decoder_max_seq_len is not None else self.max_seq_len
        dev = self.norm.scale.device
        L, kvd = len(self.layers), self.num_kv_heads * self.head_dim
        self._kc = torch.zeros(L, max_batch_size, n, kvd, dtype=BF16, device=dev)
        self._vc = torch.zeros(L, max_batch_size, n, kvd, dtype=BF16, device=dev)
        self._cache_len = 0

    def reset_caches(self):
        if self.caches_are_enabled():
            self._cache_len = 0

    @torch.no_grad()
    def infer(self, x: torch.Tensor, pos0: int) -> torch.Tensor:
        """Cached forward of S new positions pos0 .. pos0+S-1 (same for every sample): bf16 [B,S,D] -> bf16 [B,S,D].
        Same kernels as training; K / V of the new positions are appended to the caches.  pos0 == 0 (prefill) runs the
        causal attention kernels on the new block, S == 1 the decode kernel over the cache."""
        if not self.caches_are_enabled():
            raise RuntimeError("backbone caches are not enabled")          # model.py:164
        B, S, D = x.shape
        if pos0 != self._cache_len:
            raise RuntimeError(f"KV cache holds {self._cache_len} positions, cannot continue at position {pos0}")
        if B > self._kc.shape[1] or pos0 + S > self._kc.shape[2]:
            raise RuntimeError(f"KV cache too small for batch {B}, positions up to {pos0 + S}")
        if pos0 > 0 and S != 1:
            raise NotImplementedError("cached forward: several new positions behind a non-empty cache (chunked "
                                      "prefill) — generate_frame only prefills once and then steps by one")
        H, KV, hd, eps = self.num_heads, self.num_kv_heads, self.head_dim, self.norm_eps
        nq, nkv = H * hd, KV * hd
        N = B * S
        params = list(self.parameters())
        index_of = {id(p): i for i, p in enumerate(params)}
        rope = self.rope_cache(x.device)[pos0:]                              # row s of the block = position pos0 + s
        res_dtype = torch.float32 if _ag.FP32_RESIDUAL else BF16
        cur = x.reshape(N, D).contiguous()
        for li, layer in enumerate(self.layers):
            a = layer.attn
            gqkv = _ag._Group([a.q_proj, a.k_proj, a.v_proj], index_of, self._packed)
            g13 = _ag._Group([layer.mlp.w1, layer.mlp.w3], index_of, self._packed)
            lo, l2 = _ag._Lin(a.output_proj, index_of), _ag._Lin(layer.mlp.w2, index_of)
            I = layer.mlp.w1.weight.shape[0]
            xn, _ = ops.rmsnorm(cur, layer.sa_norm.scale, eps)
            qkv, _ = gqkv.fwd(xn, rope=(rope, S, nq + nkv, hd))
            q, k, v = qkv[:, :nq], qkv[:, nq:nq + nkv], qkv[:, nq + nkv:]
            self._kc[li, :B, pos0:pos0 + S].copy_(k.view(B, S, nkv))
            self._vc[li, :B, pos0:pos0 + S].copy_(v.view(B, S, nkv))
            if pos0 == 0:
                o, _ = ops.attention_fwd(q, k, v, B, S, H, KV, hd)
            else:
                o = ops.attention_decode(q, self._kc[li], self._vc[li], B, H, KV, hd, pos0 + 1)
            h, _ = lo.fwd(o, residual=cur, out_dtype=res_dtype)
            hn, _ = ops.rmsnorm(h, layer.mlp_norm.scale, eps)
            if ops.swiglu_fusable(N, I, D):
                _, act, _ = g13.fwd_swiglu(hn)
            else:
                gu, _ = g13.fwd(hn)
                act = ops.swiglu(gu[:, :I], gu[:, I:])
            cur, _ = l2.fwd(act, residual=h, out_dtype=res_dtype)
        y, _ = ops.rmsnorm(cur, self.norm.scale, eps)
        self._cache_len = pos0 + S
        return y.view(B, S, D)

    def run(self, x: torch.Tensor, adapter_rows: Optional[torch.Tensor] = None, packing=None) -> torch.Tensor:
        """bf16 [B,S,D] -> bf16 [B,S,D] (layers + final norm) through the CUDA kernels.  ``adapter_rows`` int32 [B*S]:
        which LoRA adapter each row uses (multi-adapter batching; csm/models/lora.py).  ``packing`` = (seg_start,
        seg_end, positions): several samples per row (csm/data/frames.py::pack_samples)."""
        if not x.is_cuda:
            raise RuntimeError("csm_b200: the transformer runs on CUDA only (no CPU fallback)")
        if x.dtype != BF16:
            raise RuntimeError(f"csm_b200: bf16 activations expected, got {x.dtype}; call model.to(torch.bfloat16)")
        self._adapter_rows, self._packing = adapter_rows, packing
        try:
            return StackFn.apply(x, self, *list(self.parameters()))
        finally:
            self._adapter_rows = self._packing = None

    def forward(self, h: torch.Tensor, *, input_pos: Optional[torch.Tensor] = None,
                mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        # input_pos must be arange(S) per row and mask the causal mask it indexes (utils.py:81-90): the kernels
        # implement exactly that case (is_causal) and positions row % S.
        return self.run(h).float()


def _stack(**kw) -> TransformerDecoder:
    return TransformerDecoder(**kw)


def llama3_2_1B() -> TransformerDecoder:                     # model.py:11-25
    return _stack(num_layers=16, num_heads=32, num_kv_heads=8, embed_dim=2048, max_seq_len=2048,
                  intermediate_dim=8192, norm_eps=1e-5, rope_base=500_000, scale_factor=32)


def llama3_2_100M() -> TransformerDecoder:                   # model.py:28-42
    return _stack(num_layers=4, num_heads=8, num_kv_heads=2, embed_dim=1024, max_seq_len=2048,
                  intermediate_dim=8192, norm_eps=1e-5, rope_base=500_000, scale_factor=32)


def _tiny_backbone() -> TransformerDecoder:                  # tests/create_test_model.py:42-51,81,123-131
    return _stack(num_layers=2, num_heads=4, num_kv_heads=4, embed_dim=32, max_seq_len=2048, intermediate_dim=128)


def _tiny_decoder() -> TransformerDecoder:                   # tests/create_test_model.py:179-188
    return _stack(num_layers=1, num_heads=2, num_kv_heads=2, embed_dim=16, max_seq_len=2048, intermediate_dim=64)


def _small_backbone() -> TransformerDecoder:
    return _stack(num_layers=2, num_heads=4, num_kv_heads=1, embed_dim=256, max_seq_len=2048, intermediate_dim=512)


def _small_decoder() -> TransformerDecoder:
    return _stack(num_layers=1, num_heads=2, num_kv_heads=1, embed_dim=256, max_seq_len=2048, intermediate_dim=512)


FLAVORS = {
    "llama-1B": llama3_2_1B,
    "llama-100M": llama3_2_100M,
    # test shapes (not in the reference's registry)
    "tiny-backbone": _tiny_backbone,
    "tiny-decoder": _tiny_decoder,
    "small-backbone": _small_backbone,
    "small-decoder": _small_decoder,
}


def _prepare_transformer(model):                             # model.py:51-56
    return model, model.embed_dim


def _create_causal_mask(seq_len: int, device: torch.device):  # model.py:59-61
    return torch.tril(torch.ones(seq_len, seq_len, dtype=torch.bool, device=device))


def _index_causal_mask(mask: torch.Tensor, input_pos: torch.Tensor):  # model.py:64-76
    return mask[input_pos, :]


def _multinomial_sample_one_no_sync(probs):                   # model.py:79-82
    q = torch.empty_like(probs).exponential_(1)
    return torch.argmax(probs / q, dim=-1, keepdim=True).to(dtype=torch.int)


def sample_topk(logits: torch.Tensor, topk: int, temperature: float):   # model.py:85-96
    """Top-k filter + softmax + exponential-race draw, on the device without a synchronisation (a few launches on a
    [B, 2051] row: sampling policy, not arithmetic of the model)."""
    logits = logits / temperature
    indices_to_remove = logits < torch.topk(logits, topk)[0][..., -1, None]
    scores = torch.nn.functional.log_softmax(logits.masked_fill(indices_to_remove, -float("inf")), dim=-1)
    return _multinomial_sample_one_no_sync(torch.nn.functional.softmax(scores, dim=-1))


@dataclass
class ModelArgs:                                             # model.py:99-107
    backbone_flavor: str
    decoder_flavor: str
    text_vocab_size: int
    audio_vocab_size: int
    audio_num_codebooks: int


class Model(nn.Module):
    def __init__(self, args: ModelArgs):
        super().__init__()
        self.args = args
        self.backbone, backbone_dim = _prepare_transformer(FLAVORS[args.backbone_flavor]())
        self.decoder, decoder_dim = _prepare_transformer(FLAVORS[args.decoder_flavor]())
        self.text_embeddings = nn.Embedding(args.text_vocab_size, backbone_dim)
        self.audio_embeddings = nn.Embedding(args.audio_vocab_size * args.audio_num_codebooks, backbone_dim)
        self.projection = nn.Linear(backbone_dim, decoder_dim, bias=False)
        self.codebook0_head = nn.Linear(backbone_dim, args.audio_vocab_size, bias=False)
        self.audio_head = nn.Parameter(torch.empty(args.audio_num_codebooks - 1, decoder_dim, args.audio_vocab_size))
        self._head_t = None
        self._head_t_version = -1

    # ---- reference helpers kept for API parity -------------------------------------------------
    def setup_caches(self, max_batch_size: int) -> None:
        """model.py:128-138: KV caches of both stacks (inference only; the training forward ignores them) and the
        causal-mask buffers the reference's compute_loss indexes."""
        device = next(self.parameters()).device
        if device.type == "cuda":
            self.backbone.setup_caches(max_batch_size, BF16)
            self.decoder.setup_caches(max_batch_size, BF16, decoder_max_seq_len=self.args.audio_num_codebooks)
        self.register_buffer("backbone_causal_mask", _create_causal_mask(self.backbone.max_seq_len, device),
                             persistent=False)
        self.register_buffer("decoder_causal_mask", _create_causal_mask(self.args.audio_num_codebooks, device),
                             persistent=False)

    def reset_caches(self):                                                          # model.py:197-200
        self.backbone.reset_caches()
        self.decoder.reset_caches()

    def _index_causal_mask(self, mask: torch.Tensor, input_pos: torch.Tensor) -> torch.Tensor:
        """Method form expected by compute_loss (utils.py:90) — a module-level function in the reference."""
        return _index_causal_mask(mask, input_pos)

    def _embed_audio(self, codebook: int, tokens: torch.Tensor) -> torch.Tensor:     # model.py:202-204
        return self.audio_embeddings(tokens + codebook * self.args.audio_vocab_size)

    def _embed_tokens(self, tokens: torch.Tensor) -> torch.Tensor:                   # model.py:206-217
        """[B,S,33] -> [B,S,33,D], materialised.  API parity only (torch indexing = data movement); the training
        forward never materialises this tensor: it uses the fused gather-sum kernel."""
        C, V = self.args.audio_num_codebooks, self.args.audio_vocab_size
        text = self.text_embeddings(tokens[:, :, -1]).unsqueeze(-2)
        idx = tokens[:, :, :-1] + V * torch.arange(C, device=tokens.device)
        audio = self.audio_embeddings(idx.view(-1)).reshape(tokens.size(0), tokens.size(1), C, -1)
        return torch.cat([audio, text], dim=-2)

    @torch.no_grad()
    def generate_frame(self, tokens: torch.Tensor, tokens_mask: torch.Tensor, input_pos: torch.Tensor,
                       temperature: float, topk: int, *, return_logits: bool = False,
                       forced_codes: Optional[torch.Tensor] = None):
        """model.py:140-195 on the training kernels with KV caches: tokens [B,S,33], tokens_mask
        [B,S,33], input_pos [B,S] (pos0 + arange(S), the same for every sample) -> sampled codes int32 [B,32].
        Backbone: fused gather-sum, cached forward (prefill: causal attention kernels; later frames: the decode kernel),
        codebook0 head GEMM; depth decoder: 31 cached steps (positions 0,1 together, then one by one) with a fresh cache
        per frame.  Test aids: ``return_logits`` also returns the 32 logit rows, ``forced_codes`` int [B,32] replaces
        the sampled codes that are fed back (teacher forcing), so two implementations can be compared step by step."""
        if not tokens.is_cuda:
            raise RuntimeError("csm_b200: generate_frame needs CUDA tensors (no CPU fallback)")
        assert self.backbone.caches_are_enabled(), "backbone caches are not enabled"      # model.py:164
        B, S, _ = tokens.shape
        C, V = self.args.audio_num_codebooks, self.args.audio_vocab_size
        pos0 = int(input_pos[0, 0])
        if not torch.equal(input_pos, (pos0 + torch.arange(S, device=input_pos.device)).expand(B, S)):
            raise NotImplementedError("generate_frame: input_pos must be pos0 + arange(S), identical for every sample")
        h0 = ops.embed_gather_sum(tokens, tokens_mask, self.audio_embeddings.weight, self.text_embeddings.weight)
        h = self.backbone.infer(h0, pos0)
        last_h = h[:, -1, :].contiguous()
        logits = [ops.gemm(last_h, self.codebook0_head.weight)]
        sample = sample_topk(logits[0], topk, temperature)
        if forced_codes is not None:
            sample = forced_codes[:, 0:1].to(sample.dtype)
        out = [sample]
        emb = self._embed_audio(0, sample.long())                                     # [B,1,D]
        x = torch.cat([last_h.unsqueeze(1), emb], dim=1)                             # positions 0, 1
        self.decoder.reset_caches()                                                   # model.py:180-181
        head_t = self._audio_head_t()
        dpos = 0
        for i in range(1, C):
            n_new = x.shape[1]
            xp = ops.gemm(x.reshape(B * n_new, -1).contiguous(), self.projection.weight)
            y = self.decoder.infer(xp.view(B, n_new, -1), dpos)
            dpos += n_new
            logits.append(ops.gemm(y[:, -1, :].contiguous(), head_t[i - 1]))          # == mm(h, audio_head[i-1])
            sample = sample_topk(logits[-1], topk, temperature)
            if forced_codes is not None:
                sample = forced_codes[:, i:i + 1].to(sample.dtype)
            out.append(sample)
            x = self._embed_audio(i, sample.long())
        codes = torch.cat(out, dim=1)
        return (codes, logits) if return_logits else codes

    # ---- B200 training forward ---------------------------------------------------------------
    def embed(self, tokens: torch.Tensor, tokens_mask: torch.Tensor) -> torch.Tensor:
        """A2: fused gather + mask + 33-way sum -> bf16 [B,S,D]."""
        return EmbedGatherSumFn.apply(tokens, tokens_mask, self.audio_embeddings.weight, self.text_embeddings.weight,
                                      getattr(self, "_text_grad_exchange", None))

    def _audio_head_t(self) -> torch.Tensor:
        """[31, V, Dd] shadow of audio_head [31, Dd, V] (rows of V=2051 bf16 are not 16-byte aligned, so the native
        layout cannot be a TMA operand); a frozen head (LoRA) is transposed once, a trainable one every forward."""
        ah = self.audio_head
        if self._head_t is None or self._head_t.device != ah.device or self._head_t.dtype != ah.dtype:
            self._head_t = ah.detach().transpose(1, 2).contiguous()
            self._head_t_version = ah._version
        elif ah.requires_grad or self._head_t_version != ah._version:
            # a trainable head is re-read on every forward: a replayed CUDA graph (and any kernel that updates the
            # parameter through its raw pointer) does not move the version counter.  Same buffer, so the copy is
            # captured with the step.
            self._head_t.copy_(ah.detach().transpose(1, 2))
            self._head_t_version = ah._version
        return self._head_t

    @staticmethod
    def select_frames(tokens_mask: torch.Tensor, target_len: int, fraction: float = 1.0 / 16,
                      generator: Optional[torch.Generator] = None,
                      target_lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        """A8 (docs/reference/sesame_csm/training.md:58-62): per sample keep ceil(T_b/16) of its TARGET frames,
        uniformly at random; returns int64 [N_sel, 2] of (b, p).  Position p pairs the backbone state h[b,p] with the
        target row targets[b,p,:] — the pairing the reference's semantic term fixes (utils.py:98-105) — so the valid
        positions of sample b are p < min(S-1, T_b), T_b = that sample's true target length (``target_lengths``, carried
        through collate; without it the padded length ``target_len``).  Frames in the zero-padded tail of the targets
        are never selected.  Runs on the host (the data pipeline calls it before the H2D copy)."""
        B, S = tokens_mask.shape[0], tokens_mask.shape[1]
        lens = [int(target_len)] * B if target_lengths is None else [int(x) for x in target_lengths.tolist()]
        out = []
        for b in range(B):
            limit = max(0, min(S - 1, int(target_len), lens[b]))
            if limit == 0:
                continue
            keep = max(1, math.ceil(limit * fraction))
            sel = torch.randperm(limit, generator=generator)[:keep].sort().values
            out.append(torch.stack([torch.full_like(sel, b), sel], dim=1))
        return torch.cat(out, 0) if out else torch.zeros(0, 2, dtype=torch.int64)

    def forward(self, tokens: torch.Tensor, tokens_mask: torch.Tensor,
                target_audio_tokens: Optional[torch.Tensor] = None, *, frame_idx: Optional[torch.Tensor] = None,
                decoder_frame_fraction: float = 1.0 / 16, semantic_weight: float = 100.0,
                acoustic_weight: float = 1.0, target_lengths: Optional[torch.Tensor] = None,
                mask_padded_targets: bool = False, speaker_ids: Optional[torch.Tensor] = None,
                segment_starts: Optional[torch.Tensor] = None, segment_ends: Optional[torch.Tensor] = None,
                target_mask: Optional[torch.Tensor] = None):
        """tokens int64 [B,S,33], tokens_mask bool [B,S,33], target_audio_tokens int64 [B,T,32].
        (Or the compact device format of csm/data/frames.py::pack_tokens in their place: tokens int32 [B,S,33] of
        pre-offset table rows and tokens_mask int64 [B,S], one mask word per frame — bit-identical results.)

        Returns (loss, {"semantic_loss", "acoustic_loss", "per_codebook_loss": fp32[32]}); with
        target_audio_tokens=None returns the backbone hidden state bf16 [B,S,D].

        ``target_lengths`` int64 [B] (true target frames per sample before collate's zero padding): restricts the
        decoder frame selection to real target frames, and with ``mask_padded_targets`` also drops the padded rows from
        the semantic term (the reference averages over them, utils.py:101-105: the default keeps that).
        ``speaker_ids`` int [B] (multi-adapter LoRA models only): sample b runs through adapter speaker_ids[b] of every
        adapted projection (index into the adapters, not the user's speaker number); a stack whose adapters are shared
        (one adapter) ignores it.
        ``segment_starts`` / ``segment_ends`` int32 [B,S] (sequence packing, csm/data/frames.py::pack_samples): row b
        holds several samples back to back; position i belongs to the sample spanning [start, end).  Attention is then
        block-diagonal causal, RoPE positions restart per sample, and the semantic term covers ``target_mask`` [B,S]
        (default: every position that is not the last of its sample) — i.e. exactly the positions the reference's
        ``[:, :-1]`` rule (utils.py:98-104) keeps when each sample is run alone, without any padding frame.
        """
        if not tokens.is_cuda:
            raise RuntimeError("csm_b200: Model.forward needs CUDA tensors (no CPU fallback)")
        if self.codebook0_head.weight.dtype != BF16:
            raise RuntimeError("csm_b200: the kernels compute in bf16; call model.to(torch.bfloat16) first")
        B, S, W = tokens.shape
        C, V = self.args.audio_num_codebooks, self.args.audio_vocab_size
        h0 = self.embed(tokens, tokens_mask)
        rows_b = rows_d = None
        if speaker_ids is not None:
            sid = speaker_ids.to(device=tokens.device, dtype=torch.int32)
            if getattr(self.backbone, "lora_adapters", 1) > 1:
                rows_b = sid.repeat_interleave(S).contiguous()
        elif max(getattr(self.backbone, "lora_adapters", 1), getattr(self.decoder, "lora_adapters", 1)) > 1:
            raise RuntimeError("this model holds several LoRA adapters per projection: pass speaker_ids [B]")
        packing = None
        if segment_starts is not None:
            if segment_ends is None:
                raise RuntimeError("sequence packing needs segment_starts and segment_ends")
            ss = segment_starts.to(device=tokens.device, dtype=torch.int32).contiguous()
            se = segment_ends.to(device=tokens.device, dtype=torch.int32).contiguous()
            pos = (torch.arange(S, device=tokens.device, dtype=torch.int32)[None, :] - ss).reshape(-1).contiguous()
            packing = (ss, se, pos)
        hb = self.backbone.run(h0, rows_b, packing)              # [B,S,D] bf16 (final norm applied)
        if target_audio_tokens is None:
            return hb
        T = target_audio_tokens.shape[1]
        if T < S - 1 and segment_starts is None:
            raise RuntimeError(f"target_audio_tokens has {T} frames, need at least seq_len-1 = {S - 1}")
        # ---- semantic term (utils.py:98-107): position p predicts targets[b,p,0], p < S-1, mean over B*(S-1)
        tgt0 = torch.full((B, S), -1, dtype=torch.int64, device=tokens.device)
        tgt0[:, : S - 1] = target_audio_tokens[:, : S - 1, 0]
        count = B * (S - 1)
        if packing is not None:
            idx = torch.arange(S, device=tokens.device)[None, :]
            keep = (idx + 1 < packing[1]) if target_mask is None else target_mask.to(tokens.device).bool()
            keep = keep & (idx < T)
            tg = target_audio_tokens[:, :S, 0]
            if T < S:
                tg = torch.nn.functional.pad(tg, (0, S - T))
            tgt0 = torch.where(keep, tg, torch.full_like(tg, -1))
            count = keep.sum().clamp(min=1).to(torch.float32)
        elif mask_padded_targets and target_lengths is not None:
            tl = target_lengths.to(tokens.device).clamp(max=S - 1)
            tgt0.masked_fill_(torch.arange(S, device=tokens.device)[None, :] >= tl[:, None], -1)
            count = tl.sum().clamp(min=1).to(torch.float32)          # device scalar: graph-replayable
        sem, _ = LinearCEFn.apply(hb.view(B * S, -1), self.codebook0_head.weight, tgt0.view(-1), count)
        per_cb = [sem.detach()]
        # ---- acoustic term (A7/A8)
        if frame_idx is None:
            if packing is not None:
                raise RuntimeError("a packed batch carries its own frame_idx (csm/data/frames.py::pack_samples picks the "
                                   "decoder frames per sample)")
            frame_idx = self.select_frames(tokens_mask, T, decoder_frame_fraction, target_lengths=target_lengths)
        frame_idx = frame_idx.to(tokens.device)
        if frame_idx.numel() > 0:
            Ns = frame_idx.shape[0]
            x = DecoderInputFn.apply(hb, self.audio_embeddings.weight, target_audio_tokens, frame_idx, C, V)
            xp = LinearFn.apply(x.view(Ns * C, -1), self.projection.weight)
            if speaker_ids is not None and getattr(self.decoder, "lora_adapters", 1) > 1:
                rows_d = sid[frame_idx[:, 0]].repeat_interleave(C).contiguous()
            y = self.decoder.run(xp.view(Ns, C, -1), rows_d)
            codes = target_audio_tokens[frame_idx[:, 0], frame_idx[:, 1]].contiguous()      # [Ns, C] int64 gather
            ac, rows = GroupedLinearCEFn.apply(y, self.audio_head, self._audio_head_t(), codes)
            per_cb_ac = rows.mean(dim=1)
        else:
            ac = torch.zeros((), dtype=torch.float32, device=tokens.device)
            per_cb_ac = torch.zeros(C - 1, dtype=torch.float32, device=tokens.device)
        loss = semantic_weight * sem + acoustic_weight * ac
        details: Dict[str, torch.Tensor] = {
            "semantic_loss": sem, "acoustic_loss": ac,
            "per_codebook_loss": torch.cat([per_cb[0].reshape(1), per_cb_ac.detach()])}
        return loss, details
