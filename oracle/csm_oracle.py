"""ORACLE — test infrastructure only.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs may import this package; the
product (csm-train-pytorch_b200/) never does and fails loudly without its CUDA
library.

Plain-PyTorch CPU restatement of the CSM training step:

  * ``OracleModel``      restates ``Model`` (/root/reference/src/csm/models/model.py:110-217):
                         same parameters / state-dict keys, ``_embed_tokens`` (:206-217),
                         ``_embed_audio`` (:202-204).
  * ``oracle_forward``   = reference ``compute_loss`` semantic term
                         (/root/reference/src/csm/training/utils.py:78-107) + the teacher-forced
                         restatement of ``generate_frame`` (model.py:171-193) for the acoustic
                         term that the reference leaves as a placeholder (utils.py:109-117),
                         on an explicit ``frame_idx`` subsample (docs/reference/sesame_csm/
                         training.md:52-68).
  * ``apply_lora``       LoRA math of /root/reference/src/csm/mlx/components/lora.py:71-105.
  * ``synthetic_batch``  seeded inputs of SURVEY.md §8(d).

PARITY STATUS: the reference holds no golden vector / known-answer test for this path
 => "parity unpinned" by the reference's own tests.  The restatement is
pinned instead against the reference's own code run here: tests/golden/make_golden.py
imports /root/reference's model.py + compute_loss verbatim (through oracle/torchtune_shim.py)
and checks this file bit-for-bit on the semantic path before writing tests/golden/*.pt.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import torchtune_shim as tt


@dataclass
class StackCfg:
    num_layers: int
    num_heads: int
    num_kv_heads: int
    embed_dim: int
    intermediate_dim: int
    max_seq_len: int = 2048
    norm_eps: float = 1e-5
    rope_base: float = 500_000.0
    scale_factor: float = 32.0


@dataclass
class OracleCfg:
    backbone: StackCfg
    decoder: StackCfg
    text_vocab_size: int
    audio_vocab_size: int
    audio_num_codebooks: int


def cfg_csm_1b(max_seq_len: int = 2048) -> OracleCfg:
    """model.py:11-42 (llama3_2_1B / llama3_2_100M), trainer.py:100-106."""
    return OracleCfg(
        backbone=StackCfg(16, 32, 8, 2048, 8192, max_seq_len),
        decoder=StackCfg(4, 8, 2, 1024, 8192, max_seq_len),
        text_vocab_size=128_256, audio_vocab_size=2051, audio_num_codebooks=32)


def cfg_tiny() -> OracleCfg:
    """Shapes of /root/reference/tests/create_test_model.py:42-51,81,123-131,179-188
    (hidden 32, 2 layers, 4 heads kv=heads, MLP 4x, text vocab 1000, audio vocab 200,
    decoder hidden 16 / 1 layer / 2 heads) with the real 32 codebooks (BASELINE config 1)."""
    return OracleCfg(
        backbone=StackCfg(2, 4, 4, 32, 128, 2048),
        decoder=StackCfg(1, 2, 2, 16, 64, 2048),
        text_vocab_size=1000, audio_vocab_size=200, audio_num_codebooks=32)


def cfg_small() -> OracleCfg:
    """A tensor-core-shaped miniature (head_dim 64 / 128, GQA 4:1 like CSM-1B) that the CPU
    oracle finishes in seconds; used by the GPU parity tests for the tcgen05 paths."""
    return OracleCfg(
        backbone=StackCfg(2, 4, 1, 256, 512, 2048),
        decoder=StackCfg(1, 2, 1, 256, 512, 2048),
        text_vocab_size=1000, audio_vocab_size=2051, audio_num_codebooks=32)


CONFIGS = {"csm-1b": cfg_csm_1b, "tiny": cfg_tiny, "small": cfg_small}


def _stack(c: StackCfg) -> tt.TransformerDecoder:
    m = tt.llama3_2(vocab_size=8, num_layers=c.num_layers, num_heads=c.num_heads,
                    num_kv_heads=c.num_kv_heads, embed_dim=c.embed_dim, max_seq_len=c.max_seq_len,
                    intermediate_dim=c.intermediate_dim, attn_dropout=0.0, norm_eps=c.norm_eps,
                    rope_base=c.rope_base, scale_factor=c.scale_factor)
    m.tok_embeddings = nn.Identity()   # model.py:51-56
    m.output = nn.Identity()
    return m


class OracleModel(nn.Module):
    """Restates model.py:110-126 (parameters) and :202-217 (embedding helpers)."""

    def __init__(self, cfg: OracleCfg):
        super().__init__()
        self.cfg = cfg
        self.backbone = _stack(cfg.backbone)
        self.decoder = _stack(cfg.decoder)
        D, Dd = cfg.backbone.embed_dim, cfg.decoder.embed_dim
        self.text_embeddings = nn.Embedding(cfg.text_vocab_size, D)
        self.audio_embeddings = nn.Embedding(cfg.audio_vocab_size * cfg.audio_num_codebooks, D)
        self.projection = nn.Linear(D, Dd, bias=False)
        self.codebook0_head = nn.Linear(D, cfg.audio_vocab_size, bias=False)
        self.audio_head = nn.Parameter(torch.empty(cfg.audio_num_codebooks - 1, Dd, cfg.audio_vocab_size))

    def _embed_audio(self, codebook: int, tokens: torch.Tensor) -> torch.Tensor:      # model.py:202-204
        return self.audio_embeddings(tokens + codebook * self.cfg.audio_vocab_size)

    def _embed_tokens(self, tokens: torch.Tensor) -> torch.Tensor:                    # model.py:206-217
        C, V = self.cfg.audio_num_codebooks, self.cfg.audio_vocab_size
        text = self.text_embeddings(tokens[:, :, -1]).unsqueeze(-2)
        idx = tokens[:, :, :-1] + V * torch.arange(C, device=tokens.device)
        audio = self.audio_embeddings(idx.view(-1)).reshape(tokens.size(0), tokens.size(1), C, -1)
        return torch.cat([audio, text], dim=-2)


    # ---- inference (restates model.py:128-200; pinned against the reference's own generate_frame, run live through
    # the shim, in tests/test_oracle.py::test_oracle_generate_frame_equals_reference_live)
    def setup_caches(self, max_batch_size: int) -> None:                                  # model.py:128-138
        dtype = next(self.parameters()).dtype
        device = next(self.parameters()).device
        self.backbone.setup_caches(max_batch_size, dtype)
        self.decoder.setup_caches(max_batch_size, dtype, decoder_max_seq_len=self.cfg.audio_num_codebooks)
        self.backbone_causal_mask = torch.tril(torch.ones(self.backbone.max_seq_len, self.backbone.max_seq_len,
                                                          dtype=torch.bool, device=device))
        self.decoder_causal_mask = torch.tril(torch.ones(self.cfg.audio_num_codebooks, self.cfg.audio_num_codebooks,
                                                         dtype=torch.bool, device=device))

    def reset_caches(self):                                                               # model.py:197-200
        self.backbone.reset_caches()
        self.decoder.reset_caches()

    @torch.no_grad()
    def generate_frame(self, tokens, tokens_mask, input_pos, temperature: float, topk: int, return_logits=False):
        """model.py:140-195.  topk == 1 makes sample_topk the argmax (every other probability is exactly 0), which is
        what the parity tests use.  ``return_logits``: also the 32 logit rows that were sampled from (test aid)."""
        dtype = next(self.parameters()).dtype
        assert self.backbone.caches_are_enabled(), "backbone caches are not enabled"
        curr_backbone_mask = self.backbone_causal_mask[input_pos, :]
        h = (self._embed_tokens(tokens) * tokens_mask.unsqueeze(-1)).sum(dim=2)
        h = self.backbone(h, input_pos=input_pos, mask=curr_backbone_mask).to(dtype=dtype)
        last_h = h[:, -1, :]
        c0_logits = self.codebook0_head(last_h)
        logits = [c0_logits]
        c0_sample = sample_topk(c0_logits, topk, temperature)
        c0_embed = self._embed_audio(0, c0_sample)
        curr_h = torch.cat([last_h.unsqueeze(1), c0_embed], dim=1)
        curr_sample = c0_sample.clone()
        curr_pos = torch.arange(0, curr_h.size(1), device=curr_h.device).unsqueeze(0).repeat(curr_h.size(0), 1)
        self.decoder.reset_caches()
        for i in range(1, self.cfg.audio_num_codebooks):
            curr_decoder_mask = self.decoder_causal_mask[curr_pos, :]
            decoder_h = self.decoder(self.projection(curr_h), input_pos=curr_pos, mask=curr_decoder_mask).to(dtype=dtype)
            ci_logits = torch.mm(decoder_h[:, -1, :], self.audio_head[i - 1])
            logits.append(ci_logits)
            ci_sample = sample_topk(ci_logits, topk, temperature)
            curr_h = self._embed_audio(i, ci_sample)
            curr_sample = torch.cat([curr_sample, ci_sample], dim=1)
            curr_pos = curr_pos[:, -1:] + 1
        return (curr_sample, logits) if return_logits else curr_sample


def sample_topk(logits: torch.Tensor, topk: int, temperature: float) -> torch.Tensor:
    """model.py:85-96 (+ _multinomial_sample_one_no_sync :79-82): top-k filter, softmax, Gumbel-style draw."""
    logits = logits / temperature
    indices_to_remove = logits < torch.topk(logits, topk)[0][..., -1, None]
    scores = F.log_softmax(logits.masked_fill(indices_to_remove, -float("inf")), dim=-1)
    probs = F.softmax(scores, dim=-1)
    q = torch.empty_like(probs).exponential_(1)
    return torch.argmax(probs / q, dim=-1, keepdim=True).to(dtype=torch.int)


def gather_indices(tokens: torch.Tensor, audio_vocab: int, codebooks: int) -> torch.Tensor:
    """The integer half of A2: idx[b,s,c] = tokens[b,s,c] + c*V (int64), c<32;
    column 32 is the raw text token.  Bit-exact parity target for the gather kernel."""
    idx = tokens.clone()
    idx[:, :, :codebooks] += audio_vocab * torch.arange(codebooks, device=tokens.device)
    return idx


def init_weights(model: nn.Module, seed: int = 0, std: float = 0.02) -> None:
    """Linear/Embedding/audio_head ~ N(0, 0.02) (create_test_model.py:84-131),
    RMSNorm scale = 1.  Drawn in a fixed parameter order from one CPU generator in fp32 and
    then cast, so any model with the same state-dict keys gets identical values."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters(), key=lambda kv: kv[0]):
            if "lora_" in name:
                continue
            if name.endswith(".scale"):
                p.fill_(1.0)
            else:
                p.copy_((torch.randn(p.shape, generator=g, dtype=torch.float32) * std).to(p.dtype))


# ----------------------------------------------------------------------------- LoRA
class LoRALinear(nn.Module):
    """y = x W0^T + (alpha/r) (x A^T) B^T  — lora.py:82-105; A [r,in], B [out,r]."""

    def __init__(self, base: nn.Linear, r: int, alpha: float, use_bias: bool = False):
        super().__init__()
        self.weight = base.weight
        self.weight.requires_grad_(False)
        self.r, self.alpha, self.scaling = r, alpha, alpha / r
        self.lora_A = nn.Parameter(torch.zeros(r, base.in_features, dtype=base.weight.dtype))
        self.lora_B = nn.Parameter(torch.zeros(base.out_features, r, dtype=base.weight.dtype))
        if use_bias:                                              # lora.py:66: optional LoRA bias, added unscaled (:101-102)
            self.lora_bias = nn.Parameter(torch.zeros(base.out_features, dtype=base.weight.dtype))
        self.keep = None          # test aid: a fixed dropout mask / (1 - p) for the low-rank path's input (lora.py:87-90)

    def forward(self, x):
        base = F.linear(x, self.weight)
        xl = x if self.keep is None else x * self.keep.to(x.dtype).view(x.shape)
        lo = F.linear(F.linear(xl, self.lora_A), self.lora_B) * self.scaling
        if hasattr(self, "lora_bias"):
            lo = lo + self.lora_bias
        return base + lo


_TARGETS = {"q_proj": ("attn", "q_proj"), "k_proj": ("attn", "k_proj"), "v_proj": ("attn", "v_proj"),
            "o_proj": ("attn", "output_proj"), "gate_proj": ("mlp", "w1"), "up_proj": ("mlp", "w3"),
            "down_proj": ("mlp", "w2")}


def apply_lora(model: nn.Module, r: int = 8, alpha: float = 16.0,
               target_modules: Optional[Sequence[str]] = None, seed: int = 1,
               b_std: float = 0.02, use_bias: bool = False) -> List[str]:
    """lora.py:741-827 semantics: adapters on the chosen projections of backbone AND decoder;
    everything else frozen.  A ~ N(0, 1/sqrt(in)) (lora.py:62-65); B ~ N(0, b_std) — the reference
    initialises B=0 (lora.py:66), which makes every dA exactly 0, so parity runs use b_std>0
    (SURVEY §8c)."""
    target_modules = list(target_modules or ["q_proj", "v_proj"])
    for p in model.parameters():
        p.requires_grad_(False)
    g = torch.Generator(device="cpu").manual_seed(seed)
    names = []
    for stack_name in ("backbone", "decoder"):
        stack = getattr(model, stack_name)
        for li, layer in enumerate(stack.layers):
            for t in target_modules:
                parent_name, child = _TARGETS[t]
                parent = getattr(layer, parent_name)
                base = getattr(parent, child)
                lin = LoRALinear(base, r, alpha, use_bias)
                with torch.no_grad():
                    lin.lora_A.copy_((torch.randn(lin.lora_A.shape, generator=g) /
                                      math.sqrt(base.in_features)).to(lin.lora_A.dtype))
                    lin.lora_B.copy_((torch.randn(lin.lora_B.shape, generator=g) * b_std).to(lin.lora_B.dtype))
                    if use_bias:                                  # non-zero so that the parity run exercises it
                        lin.lora_bias.copy_((torch.randn(lin.lora_bias.shape, generator=g) * b_std).to(lin.lora_bias.dtype))
                setattr(parent, child, lin)
                names.append(f"{stack_name}.layers.{li}.{parent_name}.{child}")
    return names


class MultiLoRALinear(nn.Module):
    """K adapters side by side on one projection, each ROW of the input using its own (multi-speaker batching; the
    reference trains one model copy per speaker instead, multi_speaker_lora.py:276-300,378-438 — same arithmetic per
    sample):  y_n = x_n W0^T + (alpha/r) (x_n A_k^T) B_k^T  with k = adapter of row n.
    Parameter layout == the product's: lora_A [K*r, in] (rows k*r..(k+1)*r = adapter k), lora_B [out, K*r]."""

    def __init__(self, base: nn.Linear, r: int, alpha: float, K: int):
        super().__init__()
        self.weight = base.weight
        self.weight.requires_grad_(False)
        self.r, self.K, self.scaling = r, K, alpha / r
        self.lora_A = nn.Parameter(torch.zeros(K * r, base.in_features, dtype=base.weight.dtype))
        self.lora_B = nn.Parameter(torch.zeros(base.out_features, K * r, dtype=base.weight.dtype))
        self.rows = None            # int tensor broadcastable to x.shape[:-1]: adapter index of every row

    def forward(self, x):
        t = F.linear(x, self.lora_A)
        block = torch.arange(self.K * self.r, device=x.device) // self.r
        keep = (block == self.rows.to(x.device).unsqueeze(-1)).to(t.dtype)
        return F.linear(x, self.weight) + F.linear(t * keep, self.lora_B) * self.scaling


def apply_multi_lora(model: nn.Module, r: int, alpha: float, num_adapters: Dict[str, int],
                     target_modules: Optional[Sequence[str]] = None, seed: int = 1, b_std: float = 0.02) -> None:
    """``num_adapters`` = {"backbone": Kb, "decoder": Kd}; a stack with K = 1 gets the plain LoRALinear."""
    target_modules = list(target_modules or ["q_proj", "v_proj"])
    for p in model.parameters():
        p.requires_grad_(False)
    g = torch.Generator(device="cpu").manual_seed(seed)
    for stack_name in ("backbone", "decoder"):
        K = int(num_adapters.get(stack_name, 1))
        for layer in getattr(model, stack_name).layers:
            for t in target_modules:
                parent_name, child = _TARGETS[t]
                parent = getattr(layer, parent_name)
                base = getattr(parent, child)
                lin = MultiLoRALinear(base, r, alpha, K) if K > 1 else LoRALinear(base, r, alpha)
                with torch.no_grad():
                    lin.lora_A.copy_((torch.randn(lin.lora_A.shape, generator=g) /
                                      math.sqrt(base.in_features)).to(lin.lora_A.dtype))
                    lin.lora_B.copy_((torch.randn(lin.lora_B.shape, generator=g) * b_std).to(lin.lora_B.dtype))
                setattr(parent, child, lin)


def set_adapter_rows(model: nn.Module, speaker_ids: torch.Tensor, frame_idx: torch.Tensor, S: int, C: int) -> None:
    """speaker_ids int [B] (adapter index per sample): backbone rows (b, s) and decoder rows (frame, position) get
    the adapter of their sample."""
    for stack_name in ("backbone", "decoder"):
        rows = speaker_ids.view(-1, 1).expand(-1, S) if stack_name == "backbone" else \
            speaker_ids[frame_idx[:, 0]].view(-1, 1).expand(-1, C)
        for m in getattr(model, stack_name).modules():
            if isinstance(m, MultiLoRALinear):
                m.rows = rows


# ----------------------------------------------------------------------------- inputs
def synthetic_batch(cfg: OracleCfg, B: int, S: int, seed: int = 1234, fraction: float = 1 / 16,
                    s_text: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """SURVEY §8(d) synthetic inputs: per sample [text frames | audio frames | padding frames]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    C, V, Vt = cfg.audio_num_codebooks, cfg.audio_vocab_size, cfg.text_vocab_size
    s_text = min(64, max(1, S // 4)) if s_text is None else s_text
    s_pad = max(1, S // 16)
    a0, a1 = s_text, S - s_pad
    tokens = torch.zeros(B, S, C + 1, dtype=torch.int64)
    mask = torch.zeros(B, S, C + 1, dtype=torch.bool)
    tokens[:, :a0, C] = torch.randint(0, Vt, (B, a0), generator=g)
    mask[:, :a0, C] = True
    tokens[:, a0:a1, :C] = torch.randint(0, V, (B, a1 - a0, C), generator=g)
    mask[:, a0:a1, :C] = True
    targets = torch.randint(0, V, (B, S, C), generator=g)
    n_keep = max(1, math.ceil((a1 - a0) * fraction))
    fi = []
    for b in range(B):
        perm = torch.randperm(a1 - a0, generator=g)[:n_keep].sort().values + a0
        fi.append(torch.stack([torch.full_like(perm, b), perm], dim=1))
    frame_idx = torch.cat(fi, 0)
    return {"input_tokens": tokens, "input_masks": mask, "target_audio_tokens": targets,
            "frame_idx": frame_idx}


# ----------------------------------------------------------------------------- forward
def oracle_backbone(model, tokens, tokens_mask) -> torch.Tensor:
    """utils.py:81-91 without KV caches (training): returns fp32 [B,S,D]."""
    embeds = model._embed_tokens(tokens)
    masked = embeds * tokens_mask.unsqueeze(-1)
    h = masked.sum(dim=2)
    return model.backbone(h)          # mask=None => is_causal (same values as the indexed tril mask)


def oracle_decoder_ce(model, hf: torch.Tensor, codes: torch.Tensor) -> torch.Tensor:
    """A7 for Ns frames: hf [Ns, D] backbone states (model dtype), codes int64 [Ns, C] -> fp32 CE [Ns, C-1], column
    i-1 = loss of code i.  X = projection([hf, emb(0,c_0) .. emb(C-2,c_{C-2})]), Y = decoder(X) causal over the C
    positions, logits_i = Y[i] @ audio_head[i-1] (model.py:176-191)."""
    dtype = next(model.parameters()).dtype
    C = codes.shape[1]
    V = model.audio_head.shape[2]
    embs = [model._embed_audio(i, codes[:, i]) for i in range(C - 1)]
    x = torch.stack([hf] + embs, dim=1)                        # [Ns, C, D]
    y = model.decoder(model.projection(x)).to(dtype)           # [Ns, C, Dd]
    logits = torch.einsum("ncd,cdv->ncv", y[:, 1:], model.audio_head)   # [Ns, C-1, V]
    ce = F.cross_entropy(logits.float().reshape(-1, V), codes[:, 1:].reshape(-1), reduction="none")
    return ce.view(-1, C - 1)


def oracle_forward(model, tokens, tokens_mask, targets, frame_idx=None,
                   semantic_weight: float = 100.0, acoustic_weight: float = 1.0):
    """Returns (loss, {"semantic_loss","acoustic_loss","per_codebook_loss"[C]}).

    Semantic term: utils.py:98-107 (position p predicts targets[b,p,0], p < S-1, mean).
    Acoustic term (A7): for (b,p) in frame_idx: X = projection([h[b,p], emb(0,c0)..emb(C-2,c_{C-2})]),
    Y = decoder(X) causal over C positions, logits_i = Y[i] @ audio_head[i-1] predicts c_i (model.py:176-191).
    CE is evaluated in fp32 on logits produced in the model dtype."""
    dtype = next(model.parameters()).dtype
    C = model.cfg.audio_num_codebooks if hasattr(model, "cfg") else model.args.audio_num_codebooks
    V = model.codebook0_head.weight.shape[0]
    h32 = oracle_backbone(model, tokens, tokens_mask)          # fp32 (torchtune .float())
    h = h32.to(dtype)                                          # model.py:169
    S = tokens.size(1)
    sem_logits = model.codebook0_head(h[:, :-1])
    tgt0 = targets[:, :, 0][:, : sem_logits.size(1)]
    sem = F.cross_entropy(sem_logits.float().reshape(-1, V), tgt0.reshape(-1))
    per_cb = [sem]
    if frame_idx is not None and frame_idx.numel() > 0:
        b_i, p_i = frame_idx[:, 0], frame_idx[:, 1]
        assert int(p_i.max()) < min(S - 1, targets.size(1))
        ce = oracle_decoder_ce(model, h[b_i, p_i], targets[b_i, p_i])
        per_cb += list(ce.mean(0))
        ac = ce.mean()
    else:
        ac = torch.zeros((), dtype=torch.float32)
        per_cb += [ac] * (C - 1)
    loss = semantic_weight * sem + acoustic_weight * ac
    return loss, {"semantic_loss": sem, "acoustic_loss": ac,
                  "per_codebook_loss": torch.stack([p.detach().float() for p in per_cb])}
