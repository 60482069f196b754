"""Tensor-level wrappers over the C ABI (no autograd here; see csm/autograd.py).

PyTorch is plumbing only: it owns device memory and the stream; every arithmetic op on the training
path is a kernel in libcsm_b200.so.  All tensors must be CUDA, contiguous in the documented layout and
bf16 unless stated otherwise.
"""
from __future__ import annotations

import math
import os as _os
from typing import Optional, Tuple

import torch

from . import _lib

BF16 = torch.bfloat16
GEMM_AUTO, GEMM_SIMT, GEMM_TC = 0, 1, 2
_backend_override = GEMM_AUTO


def set_gemm_backend(b: int) -> None:
    """Test hook: force the scalar (1) or tcgen05 (2) GEMM back-end; 0 = automatic."""
    global _backend_override
    _backend_override = b


_gemm_prof = None   # when a list: (flops, start_event, end_event) per GEMM launch (bench.py roofline leg)


def gemm_profile_start() -> None:
    global _gemm_prof
    _gemm_prof = []


def gemm_profile_stop():
    """-> (total flops, total ms, launches) over the GEMM launches since gemm_profile_start(); caller synchronises."""
    global _gemm_prof
    rec, _gemm_prof = _gemm_prof or [], None
    return (sum(f for f, _, _ in rec), sum(a.elapsed_time(b) for _, a, b in rec), len(rec))


def set_gemm_cta_pair_mode(m: int) -> None:
    """Test hook: CTA-pair (cta_group::2) GEMM mode: -1 automatic, 0 never, 1 whenever the shape allows."""
    _lib.load().csm_set_gemm_cta_pair_mode(m)


def set_gemm_dynamic_tiles(m: int) -> None:
    """1: GEMM tiles drawn from a global counter (data-parallel runs: NCCL shares the SMs); 0 (default): static."""
    _lib.load().csm_set_gemm_dynamic_tiles(m)


_attn_backend = 0


def set_attn_backend(b: int) -> None:
    """Test hook: 0 auto, 1 scalar, 2 mma.sync, 3 tcgen05, 4 short-sequence (seq <= 32)."""
    global _attn_backend
    _attn_backend = b
    _lib.load().csm_set_attn_backend(b)


def set_pdl(on: int) -> None:
    """Programmatic dependent launch between the step's kernels (csrc/common.cuh): 0 (default) / 1 (or CSM_PDL=1)."""
    _lib.load().csm_set_pdl(int(on))


def set_attn_fwd_variant(v: int) -> None:
    """A/B hook: tcgen05 attention forward with the output accumulated in TMEM (1, default) or in registers (0)."""
    _lib.load().csm_set_attn_fwd_variant(v)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _st():
    return torch.cuda.current_stream().cuda_stream


def _chk_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("csm ops need CUDA tensors: libcsm_b200 has no CPU fallback")


# ----------------------------------------------------------------------------- embeddings
def embed_gather_sum(tokens, mask, audio_emb, text_emb, *, debug: bool = False):
    """tokens int64 [B,S,C+1], mask bool [B,S,C+1] -> h bf16 [B,S,D] (+ (idx, eff_mask, status) when debug)."""
    _chk_cuda(tokens, mask, audio_emb, text_emb)
    B, S, W = tokens.shape
    C = W - 1
    D = audio_emb.shape[1]
    V = audio_emb.shape[0] // C
    tokens = tokens.contiguous()
    mask_u8 = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8).contiguous()
    h = torch.empty(B, S, D, dtype=BF16, device=tokens.device)
    idx = msk = status = None
    if debug:
        idx = torch.empty_like(tokens)
        msk = torch.empty(B, S, W, dtype=torch.uint8, device=tokens.device)
        status = torch.zeros(1, dtype=torch.int32, device=tokens.device)
    lib = _lib.load()
    _lib.check(lib.csm_embed_gather_sum_fwd(_p(tokens), _p(mask_u8), _p(audio_emb), _p(text_emb), _p(h), _p(idx),
                                            _p(msk), _p(status), B * S, C, V, text_emb.shape[0], D, _st()),
               "embed_gather_sum_fwd")
    return (h, idx, msk, status) if debug else h


def embed_gather_sum_bwd(tokens, mask, dh, d_audio: Optional[torch.Tensor], d_text: Optional[torch.Tensor],
                         audio_vocab: int, text_vocab: int) -> None:
    _chk_cuda(tokens, mask, dh, d_audio, d_text)
    B, S, W = tokens.shape
    mask_u8 = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8).contiguous()
    lib = _lib.load()
    _lib.check(lib.csm_embed_gather_sum_bwd(_p(tokens.contiguous()), _p(mask_u8), _p(dh.contiguous()), _p(d_audio),
                                            _p(d_text), B * S, W - 1, audio_vocab, text_vocab, dh.shape[-1], _st()),
               "embed_gather_sum_bwd")


def is_packed_tokens(tokens, mask) -> bool:
    """True for the compact device format of csm/data/frames.py::pack_tokens: int32 pre-offset rows [B,S,C+1] and one
    int64 mask word per frame [B,S]."""
    return tokens.dtype == torch.int32 and mask.dtype == torch.int64 and mask.dim() == tokens.dim() - 1


def embed_gather_sum_packed(rows, mask_bits, audio_emb, text_emb, *, status=None):
    """rows int32 [B,S,C+1] (pre-offset table rows), mask_bits int64 [B,S] (bit c = mask of column c) -> h bf16 [B,S,D];
    bit-identical to embed_gather_sum on the unpacked batch."""
    _chk_cuda(rows, mask_bits, audio_emb, text_emb, status)
    assert rows.dtype == torch.int32 and mask_bits.dtype == torch.int64
    B, S, W = rows.shape
    C = W - 1
    D = audio_emb.shape[1]
    h = torch.empty(B, S, D, dtype=BF16, device=rows.device)
    lib = _lib.load()
    _lib.check(lib.csm_embed_gather_sum_packed_fwd(_p(rows.contiguous()), _p(mask_bits.contiguous()), _p(audio_emb),
                                                   _p(text_emb), _p(h), _p(status), B * S, C, audio_emb.shape[0] // C,
                                                   text_emb.shape[0], D, _st()), "embed_gather_sum_packed_fwd")
    return h


def embed_gather_sum_packed_bwd(rows, mask_bits, dh, d_audio, d_text, audio_vocab: int, text_vocab: int) -> None:
    _chk_cuda(rows, mask_bits, dh, d_audio, d_text)
    B, S, W = rows.shape
    lib = _lib.load()
    _lib.check(lib.csm_embed_gather_sum_packed_bwd(_p(rows.contiguous()), _p(mask_bits.contiguous()), _p(dh.contiguous()),
                                                   _p(d_audio), _p(d_text), B * S, W - 1, audio_vocab, text_vocab,
                                                   dh.shape[-1], _st()), "embed_gather_sum_packed_bwd")


def decoder_input(h, audio_emb, targets, frame_idx, codebooks: int, audio_vocab: int):
    """h bf16 [B,S,D], targets int64 [B,T,C], frame_idx int64 [Ns,2] -> x bf16 [Ns, C, D]."""
    _chk_cuda(h, audio_emb, targets, frame_idx)
    B, S, D = h.shape
    Ns = frame_idx.shape[0]
    x = torch.empty(Ns, codebooks, D, dtype=BF16, device=h.device)
    lib = _lib.load()
    _lib.check(lib.csm_decoder_input_fwd(_p(h), _p(audio_emb), _p(targets.contiguous()), _p(frame_idx.contiguous()),
                                         _p(x), Ns, S, targets.shape[1], codebooks, audio_vocab, D, _st()),
               "decoder_input_fwd")
    return x


def decoder_input_bwd(dx, targets, frame_idx, dh, d_audio, codebooks: int, audio_vocab: int) -> None:
    """dh bf16 [B,S,D] (accumulated in place), d_audio nullable (accumulated in place)."""
    _chk_cuda(dx, targets, frame_idx, dh, d_audio)
    B, S, D = dh.shape
    lib = _lib.load()
    _lib.check(lib.csm_decoder_input_bwd(_p(dx.contiguous()), _p(targets.contiguous()), _p(frame_idx.contiguous()),
                                         _p(dh), _p(d_audio), frame_idx.shape[0], S, targets.shape[1], codebooks,
                                         audio_vocab, D, _st()), "decoder_input_bwd")


# ----------------------------------------------------------------------------- norm / rope / swiglu
def rmsnorm(x, scale, eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """x bf16, or fp32 when it is the fp32 residual stream -> (y bf16, rstd fp32 [rows])."""
    _chk_cuda(x, scale)
    D = x.shape[-1]
    rows = x.numel() // D
    y = torch.empty(x.shape, dtype=BF16, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    _lib.check(lib.csm_rmsnorm_fwd(_p(x), _p(scale), _p(y), _p(rstd), rows, D, eps,
                                   1 if x.dtype == torch.float32 else 0, _st()), "rmsnorm_fwd")
    return y, rstd


def rmsnorm_bwd(dy, x, scale, rstd, dres: Optional[torch.Tensor], dscale_f32: Optional[torch.Tensor]):
    _chk_cuda(dy, x, scale, rstd, dres, dscale_f32)
    D = x.shape[-1]
    rows = x.numel() // D
    dx = torch.empty(x.shape, dtype=BF16, device=x.device)
    lib = _lib.load()
    _lib.check(lib.csm_rmsnorm_bwd(_p(dy), _p(x), _p(scale), _p(rstd), _p(dres), _p(dx), _p(dscale_f32), rows, D,
                                   1 if x.dtype == torch.float32 else 0, _st()), "rmsnorm_bwd")
    return dx


def rope_(x2d, cache, seq_len: int, heads: int, head_dim: int, inverse: bool = False, ld: Optional[int] = None,
          positions: Optional[torch.Tensor] = None):
    """In place on x2d [rows, >= heads*head_dim] (row stride ld); `positions` int32 [rows] overrides row % seq_len."""
    _chk_cuda(x2d, cache, positions)
    rows = x2d.shape[0]
    ld = x2d.stride(0) if ld is None else ld
    lib = _lib.load()
    _lib.check(lib.csm_rope(_p(x2d), _p(cache), rows, seq_len, heads, head_dim, ld, 1 if inverse else 0,
                            _p(positions), _st()), "rope")
    return x2d


def swiglu(gate, up):
    _chk_cuda(gate, up)
    rows, cols = gate.shape
    out = torch.empty(rows, cols, dtype=BF16, device=gate.device)
    lib = _lib.load()
    _lib.check(lib.csm_swiglu_fwd(_p(gate), _p(up), _p(out), rows, cols, gate.stride(0), up.stride(0), cols, _st()),
               "swiglu_fwd")
    return out


def swiglu_bwd(dout, gate, up, dgate=None, dup=None):
    _chk_cuda(dout, gate, up, dgate, dup)
    rows, cols = gate.shape
    dgate = torch.empty(rows, cols, dtype=BF16, device=gate.device) if dgate is None else dgate
    dup = torch.empty(rows, cols, dtype=BF16, device=gate.device) if dup is None else dup
    lib = _lib.load()
    _lib.check(lib.csm_swiglu_bwd(_p(dout), _p(gate), _p(up), _p(dgate), _p(dup), rows, cols, dout.stride(0),
                                  gate.stride(0), up.stride(0), dgate.stride(0), dup.stride(0), _st()), "swiglu_bwd")
    return dgate, dup


# ----------------------------------------------------------------------------- GEMM
SPLITK_ENABLED = True
_streamk_ws = {}       # device index -> zeroed scratch registered with the library (kept alive here)


def _ensure_streamk_workspace(device) -> None:
    """Registers the stream-K scratch of the CTA-pair GEMM once per process (first GEMM call, i.e. before any CUDA-graph
    capture).  The library keeps one pointer, so a process drives one device — the one-process-per-GPU model of §8(e)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx in _streamk_ws:
        return
    lib = _lib.load()
    nbytes = lib.csm_gemm_streamk_workspace_bytes()
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    lib.csm_gemm_set_streamk_workspace(buf.data_ptr(), nbytes)
    _streamk_ws[idx] = buf


def set_gemm_streamk_mode(m: int) -> None:
    """0 (default) whole tiles only; 1 stream-K schedule when the last wave of tiles would be badly filled."""
    _lib.load().csm_set_gemm_streamk_mode(m)


def set_gemm_narrow_tail_mode(m: int) -> None:
    """Experimental (default 0): 1 narrows the MMAs of a ragged last column tile to the columns that exist."""
    _lib.load().csm_set_gemm_narrow_tail_mode(m)


def _splitk_choice(M: int, N: int, K: int) -> int:
    """Number of reduction groups for a skinny GEMM (0 = run it as one GEMM).  Skinny = one output dimension <= 64 and
    at most 48 output tiles of 128 x 128; the split fills ~100-160 CTAs and keeps >= 256 reduction elements per group."""
    if not SPLITK_ENABLED or min(M, N) > 64 or K < 1024:
        return 0
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    if tiles > 48 or M * N * 4 * 16 > (64 << 20):
        return 0
    for s in (16, 8, 4, 2):
        if K % (s * 64) == 0 and K // s >= 256 and tiles * s <= 160:
            return s
    return 0



SKINNY_ENABLED = _os.environ.get("CSM_SKINNY", "1") != "0"


def set_skinny_mode(m: int) -> None:
    """A/B hook: 1 (default) tall-skinny LoRA products run on the streaming mma.sync kernels (csrc/skinny.cu), 0 = on
    the tcgen05 GEMM with a split reduction."""
    global SKINNY_ENABLED
    SKINNY_ENABLED = bool(m)


def _skinny(a, b, trans_a, trans_b, out, M, N, K, alpha) -> bool:
    """Routes out[M,N] = alpha * op(a) op(b) to csrc/skinny.cu when one output dimension is <= 64 and the other operand
    is a long stream: rowdot (t = x A^T, dts = dy B) or coldot (dB = dy^T t, dA = dts^T x).  False: not taken."""
    lib = _lib.load()
    if not trans_a:
        if N > 64 or M < 256:
            return False
        lay = 1 if trans_b else 0
        args = (_p(a), _p(b), _p(out), M, K, N, a.stride(0), b.stride(0), out.stride(0), lay)
        if not lib.csm_skinny_supported(0, *args):
            return False
        _lib.check(lib.csm_skinny_rowdot(*args, alpha, _st()), "skinny_rowdot")
        return True
    if not trans_b or K < 256:
        return False
    if N <= 64 and M > N:            # out[C, R]: X = a [rows, C], T = b [rows, R]
        args = (_p(a), _p(b), _p(out), K, M, N, a.stride(0), b.stride(0), out.stride(0), 0)
    elif M <= 64:                    # out[R, C]: X = b [rows, C], T = a [rows, R]
        args = (_p(b), _p(a), _p(out), K, N, M, b.stride(0), a.stride(0), out.stride(0), 1)
    else:
        return False
    if not lib.csm_skinny_supported(1, *args):
        return False
    _lib.check(lib.csm_skinny_coldot(*args, alpha, _st()), "skinny_coldot")
    return True


def gemm_rope(a, b, cache, seq_len: int, rope_cols: int, head_dim: int, *, a2=None, b2=None, positions=None):
    """out = a @ b^T (+ a2 @ b2^T) with RoPE applied to columns [0, rope_cols) (heads of head_dim, position = row %
    seq_len): the fused q|k|v projection.  One launch when the tcgen05 GEMM takes the shape, else gemm + rope."""
    _chk_cuda(a, b, cache, a2, b2)
    M, K = a.shape
    N = b.shape[0]
    be = _backend_override
    fused = (be != GEMM_SIMT and N % 32 == 0 and K >= 64 and M * N * K >= (1 << 21) and a.stride(1) == 1
             and b.stride(1) == 1 and a.stride(0) % 8 == 0 and b.stride(0) % 8 == 0)
    if not fused:
        out = gemm(a, b, a2=a2, b2=b2)
        rope_(out[:, :rope_cols], cache, seq_len, rope_cols // head_dim, head_dim, positions=positions)
        return out
    _ensure_streamk_workspace(a.device)
    out = torch.empty(M, N, dtype=BF16, device=a.device)
    K2 = a2.shape[1] if a2 is not None else 0
    lib = _lib.load()
    if _gemm_prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(lib.csm_gemm_bf16_rope(_p(a), _p(b), _p(out), M, N, K, a.stride(0), b.stride(0), out.stride(0), _p(a2),
                                      _p(b2), K2, a2.stride(0) if a2 is not None else 0,
                                      b2.stride(0) if b2 is not None else 0, _p(cache), seq_len, rope_cols, head_dim,
                                      _p(positions), _st()), "gemm_bf16_rope")
    if _gemm_prof is not None:
        e1.record()
        _gemm_prof.append((2.0 * M * N * (K + K2), e0, e1))
    return out


def gemm(a, b, *, trans_a: bool = False, trans_b: bool = False, out: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, accumulate: bool = False, alpha: float = 1.0,
         a2: Optional[torch.Tensor] = None, b2: Optional[torch.Tensor] = None, out_dtype=BF16,
         backend: Optional[int] = None):
    """out[M,N] (=|+=) alpha*(op(a) @ op(b)^T-style product [+ a2 @ b2]) [+ residual].

    a: [M,K] (or [K,M] when trans_a); b: [N,K] (nn.Linear layout; or [K,N] when trans_b); 2-D, unit inner stride.
    """
    _chk_cuda(a, b, out, residual, a2, b2)
    _ensure_streamk_workspace(a.device)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if trans_a else a.shape
    N, Kb = (b.shape[1], b.shape[0]) if trans_b else b.shape
    if K != Kb:
        raise RuntimeError(f"gemm: inner dimensions differ ({K} vs {Kb})")
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    be = _backend_override if backend is None else backend
    plain = be != GEMM_SIMT and a2 is None and residual is None and not accumulate and out.dtype == BF16
    if plain and SKINNY_ENABLED and _skinny(a, b, trans_a, trans_b, out, M, N, K, alpha):
        return out
    splits = _splitk_choice(M, N, K) if plain else 0
    if splits:
        lib = _lib.load()
        nbytes = lib.csm_gemm_splitk_workspace_bytes(M, N, splits)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
        if _gemm_prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.check(lib.csm_gemm_bf16_splitk(_p(a), _p(b), _p(out), M, N, K, a.stride(0), b.stride(0), out.stride(0),
                                            1 if trans_a else 0, 1 if trans_b else 0, alpha, splits, _p(ws), nbytes,
                                            _st()), "gemm_bf16_splitk")
        if _gemm_prof is not None:
            e1.record()
            _gemm_prof.append((2.0 * M * N * K, e0, e1))
        return out
    K2 = 0
    if a2 is not None:
        K2 = a2.shape[0] if trans_a else a2.shape[1]
        assert a2.stride(1) == 1 and b2.stride(1) == 1
    lib = _lib.load()
    if _gemm_prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    c_dtype = 1 if out.dtype == torch.float32 else 0
    if residual is not None and residual.dtype == torch.float32:
        if out.dtype != torch.float32:
            raise RuntimeError("gemm: an fp32 residual needs an fp32 output (the fp32 residual stream)")
        c_dtype |= 2                                   # CSM_DT_RES_F32
    _lib.check(lib.csm_gemm_bf16(_p(a), _p(b), _p(out), _p(residual), M, N, K, a.stride(0), b.stride(0),
                                 out.stride(0), residual.stride(0) if residual is not None else 0,
                                 1 if trans_a else 0, 1 if trans_b else 0, c_dtype,
                                 1 if accumulate else 0, alpha, _p(a2), _p(b2), K2,
                                 a2.stride(0) if a2 is not None else 0, b2.stride(0) if b2 is not None else 0,
                                 be, _st()), "gemm_bf16")
    if _gemm_prof is not None:
        e1.record()
        _gemm_prof.append((2.0 * M * N * (K + K2), e0, e1))
    return out


# ----------------------------------------------------------------------------- fused SwiGLU MLP GEMMs
def swiglu_fusable(M: int, inter: int, K: int) -> bool:
    """True when the w1|w3 GEMM + SwiGLU (and the w2 dgrad + SwiGLU backward) run as one tcgen05 launch each."""
    return _backend_override != GEMM_SIMT and bool(_lib.load().csm_gemm_swiglu_supported(M, inter, K))


def gemm_swiglu_fwd(x, w13, *, a2=None, b2=None):
    """x [M,K], w13 = [w1; w3] [2I,K] -> (gate_up bf16 [M,2I], act bf16 [M,I]) ; optional LoRA tail a2 [M,r], b2 [2I,r]."""
    _chk_cuda(x, w13, a2, b2)
    _ensure_streamk_workspace(x.device)
    M, K = x.shape
    inter = w13.shape[0] // 2
    gu = torch.empty(M, 2 * inter, dtype=BF16, device=x.device)
    act = torch.empty(M, inter, dtype=BF16, device=x.device)
    K2 = a2.shape[1] if a2 is not None else 0
    lib = _lib.load()
    if _gemm_prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(lib.csm_gemm_swiglu_fwd(_p(x), _p(w13), _p(gu), _p(act), M, inter, K, x.stride(0), w13.stride(0),
                                       gu.stride(0), act.stride(0), _p(a2), _p(b2), K2,
                                       a2.stride(0) if a2 is not None else 0, b2.stride(0) if b2 is not None else 0,
                                       _st()), "gemm_swiglu_fwd")
    if _gemm_prof is not None:
        e1.record()
        _gemm_prof.append((2.0 * M * 2 * inter * (K + K2), e0, e1))
    return gu, act


def gemm_swiglu_bwd(dy, w2, gu, *, a2=None, b2=None):
    """dy [M,K], w2 [K,I] (nn.Linear weight of the down projection), gu [M,2I] -> dgate_up bf16 [M,2I]."""
    _chk_cuda(dy, w2, gu, a2, b2)
    M, K = dy.shape
    inter = w2.shape[1]
    dgu = torch.empty(M, 2 * inter, dtype=BF16, device=dy.device)
    K2 = a2.shape[1] if a2 is not None else 0
    lib = _lib.load()
    if _gemm_prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(lib.csm_gemm_swiglu_bwd(_p(dy), _p(w2), _p(gu), _p(dgu), M, inter, K, dy.stride(0), w2.stride(0),
                                       gu.stride(0), dgu.stride(0), _p(a2), _p(b2), K2,
                                       a2.stride(0) if a2 is not None else 0, b2.stride(0) if b2 is not None else 0,
                                       _st()), "gemm_swiglu_bwd")
    if _gemm_prof is not None:
        e1.record()
        _gemm_prof.append((2.0 * M * inter * (K + K2), e0, e1))
    return dgu


# ----------------------------------------------------------------------------- attention
def attention_fwd(q, k, v, batch: int, seq: int, heads: int, kv_heads: int, head_dim: int, seg_start=None):
    """q [B*S, H*hd], k/v [B*S, KV*hd] (row strides free) -> (o [B*S, H*hd], lse fp32 [B,H,S]).
    `seg_start` int32 [B,S] (sequence packing): block-diagonal causal attention, see csm_attn_varlen_fwd."""
    _chk_cuda(q, k, v, seg_start)
    o = torch.empty(batch * seq, heads * head_dim, dtype=BF16, device=q.device)
    lse = torch.empty(batch, heads, seq, dtype=torch.float32, device=q.device)
    lib = _lib.load()
    if seg_start is not None:
        assert seg_start.dtype == torch.int32 and seg_start.is_contiguous() and seg_start.numel() == batch * seq
        _lib.check(lib.csm_attn_varlen_fwd(_p(q), _p(k), _p(v), _p(o), _p(lse), batch, seq, heads, kv_heads, head_dim,
                                           q.stride(0), k.stride(0), v.stride(0), o.stride(0),
                                           1.0 / math.sqrt(head_dim), _p(seg_start), _st()), "attn_varlen_fwd")
        return o, lse
    _lib.check(lib.csm_attn_causal_gqa_fwd(_p(q), _p(k), _p(v), _p(o), _p(lse), batch, seq, heads, kv_heads,
                                           head_dim, q.stride(0), k.stride(0), v.stride(0), o.stride(0),
                                           1.0 / math.sqrt(head_dim), _st()), "attn_fwd")
    return o, lse


def attention_bwd(q, k, v, o, lse, dout, batch, seq, heads, kv_heads, head_dim, dq=None, dk=None, dv=None,
                  rope_cache=None, seg_start=None, seg_end=None):
    """With `rope_cache` the returned dq / dk are gradients w.r.t. the UN-rotated projections (inverse RoPE applied):
    inside the tcgen05 kernels' store epilogues when they take the shape, else by the rope kernel afterwards."""
    _chk_cuda(q, k, v, o, lse, dout, dq, dk, dv)
    dev = q.device
    dq = torch.empty(batch * seq, heads * head_dim, dtype=BF16, device=dev) if dq is None else dq
    dk = torch.empty(batch * seq, kv_heads * head_dim, dtype=BF16, device=dev) if dk is None else dk
    dv = torch.empty(batch * seq, kv_heads * head_dim, dtype=BF16, device=dev) if dv is None else dv
    lib = _lib.load()
    nbytes = lib.csm_attn_bwd_workspace_bytes(batch, seq, heads, kv_heads, head_dim)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    if seg_start is not None:                          # sequence packing: tcgen05 kernels only (they raise otherwise)
        _chk_cuda(seg_start, seg_end)
        _lib.check(lib.csm_attn_varlen_bwd(_p(q), _p(k), _p(v), _p(o), _p(lse), _p(dout), _p(dq), _p(dk), _p(dv), batch,
                                           seq, heads, kv_heads, head_dim, q.stride(0), k.stride(0), v.stride(0),
                                           o.stride(0), dq.stride(0), dk.stride(0), dv.stride(0),
                                           1.0 / math.sqrt(head_dim), _p(rope_cache), _p(seg_start), _p(seg_end),
                                           _p(ws), nbytes, _st()), "attn_varlen_bwd")
        return dq, dk, dv
    fuse = (rope_cache is not None and head_dim == 64 and seq >= 128 and _attn_backend in (0, 3)
            and all(t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0 for t in (q, k, v, o, dq, dk, dv, dout)))
    if fuse:
        _lib.check(lib.csm_attn_causal_gqa_bwd_rope(_p(q), _p(k), _p(v), _p(o), _p(lse), _p(dout), _p(dq), _p(dk),
                                                    _p(dv), batch, seq, heads, kv_heads, head_dim, q.stride(0),
                                                    k.stride(0), v.stride(0), o.stride(0), dq.stride(0), dk.stride(0),
                                                    dv.stride(0), 1.0 / math.sqrt(head_dim), _p(rope_cache), _p(ws),
                                                    nbytes, _st()), "attn_bwd_rope")
        return dq, dk, dv
    _lib.check(lib.csm_attn_causal_gqa_bwd(_p(q), _p(k), _p(v), _p(o), _p(lse), _p(dout), _p(dq), _p(dk), _p(dv),
                                           batch, seq, heads, kv_heads, head_dim, q.stride(0), k.stride(0),
                                           v.stride(0), o.stride(0), dq.stride(0), dk.stride(0), dv.stride(0),
                                           1.0 / math.sqrt(head_dim), _p(ws), nbytes, _st()), "attn_bwd")
    if rope_cache is not None:
        rope_(dq, rope_cache, seq, heads, head_dim, inverse=True)
        rope_(dk, rope_cache, seq, kv_heads, head_dim, inverse=True)
    return dq, dk, dv


def attention_decode(q, k_cache, v_cache, batch: int, heads: int, kv_heads: int, head_dim: int, kv_len: int):
    """q [B, H*hd] (one new position per sample) vs caches [B_max, max_seq, KV*hd] (first kv_len positions of each of
    the first B samples) -> o bf16 [B, H*hd]."""
    _chk_cuda(q, k_cache, v_cache)
    assert k_cache.dim() == 3 and k_cache.stride(2) == 1 and k_cache.stride() == v_cache.stride()
    o = torch.empty(batch, heads * head_dim, dtype=BF16, device=q.device)
    lib = _lib.load()
    _lib.check(lib.csm_attn_decode(_p(q), _p(k_cache), _p(v_cache), _p(o), batch, heads, kv_heads, head_dim, kv_len,
                                   q.stride(0), o.stride(0), k_cache.stride(0), k_cache.stride(1),
                                   1.0 / math.sqrt(head_dim), _st()), "attn_decode")
    return o


# ----------------------------------------------------------------------------- fused linear + cross-entropy
def _ce_geometry(h, w, trans_w: bool, groups: int):
    if groups == 1:
        M, K = h.shape
        V = w.shape[1] if trans_w else w.shape[0]
        return M, V, K, h.stride(0), 0, w.stride(0), 0
    # grouped: h [M, G(+off), K] view with strides, w [G, V, K] or [G, K, V]
    M, G, K = h.shape
    assert G == groups and h.stride(2) == 1
    V = w.shape[2] if trans_w else w.shape[1]
    return M, V, K, h.stride(0), h.stride(1), w.stride(1), w.stride(0)


def linear_ce_fwd(h, w, targets, *, trans_w: bool = False, groups: int = 1, tgt_row_stride: int = 1,
                  tgt_group_stride: int = 0, backend: Optional[int] = None):
    """Returns (loss_rows fp32 [groups, M], lse fp32 [groups, M]).  `targets` is an int64 tensor whose element
    for (group g, row m) sits at offset m*tgt_row_stride + g*tgt_group_stride from its data pointer."""
    _chk_cuda(h, w, targets)
    _ensure_streamk_workspace(h.device)
    M, V, K, ldh, hgs, ldw, wgs = _ce_geometry(h, w, trans_w, groups)
    lib = _lib.load()
    nbytes = lib.csm_linear_ce_workspace_bytes(M, V, K, groups)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=h.device)
    loss = torch.empty(groups, M, dtype=torch.float32, device=h.device)
    lse = torch.empty(groups, M, dtype=torch.float32, device=h.device)
    be = _backend_override if backend is None else backend
    _lib.check(lib.csm_linear_ce_fwd(_p(h), _p(w), _p(targets), _p(loss), _p(lse), M, V, K, groups, ldh, hgs, ldw,
                                     wgs, 1 if trans_w else 0, tgt_row_stride, tgt_group_stride, _p(ws), nbytes, be,
                                     _st()), "linear_ce_fwd")
    return loss, lse


def linear_ce_bwd(h, w, targets, lse, grad_scale: float, *, dh: torch.Tensor, grad_scale_dev=None, dw: Optional[torch.Tensor] = None,
                  dw_accumulate: bool = False, trans_w: bool = False, groups: int = 1, tgt_row_stride: int = 1,
                  tgt_group_stride: int = 0, backend: Optional[int] = None):
    _chk_cuda(h, w, targets, lse, dh, dw, grad_scale_dev)
    _ensure_streamk_workspace(h.device)
    M, V, K, ldh, hgs, ldw, wgs = _ce_geometry(h, w, trans_w, groups)
    lib = _lib.load()
    nbytes = lib.csm_linear_ce_workspace_bytes(M, V, K, groups)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=h.device)
    if groups == 1:
        lddh, dhgs = dh.stride(0), 0
    else:
        lddh, dhgs = dh.stride(0), dh.stride(1)
    be = _backend_override if backend is None else backend
    _lib.check(lib.csm_linear_ce_bwd(_p(h), _p(w), _p(targets), _p(lse), grad_scale, _p(grad_scale_dev), _p(dh), _p(dw),
                                     1 if dw_accumulate else 0, M, V, K, groups, ldh, hgs, ldw, wgs,
                                     1 if trans_w else 0, tgt_row_stride, tgt_group_stride, lddh, dhgs, _p(ws),
                                     nbytes, be, _st()), "linear_ce_bwd")
    return dh, dw


# ----------------------------------------------------------------------------- multi-adapter LoRA
def lora_mask_rows_(t, adapter_ids, rank: int, adapters: int):
    """In place: row i of t [rows, cols] keeps only the `rank`-column blocks of adapter adapter_ids[i] (int32)."""
    _chk_cuda(t, adapter_ids)
    assert t.dim() == 2 and t.stride(1) == 1 and adapter_ids.dtype == torch.int32 and adapter_ids.numel() == t.shape[0]
    lib = _lib.load()
    _lib.check(lib.csm_lora_mask_rows(_p(t), t.stride(0), t.shape[0], t.shape[1], _p(adapter_ids), rank, adapters,
                                      _st()), "lora_mask_rows")
    return t


def lora_dropout(x, p: float, seed_dev, salt: int, out=None, accumulate: bool = False):
    """out (=|+=) x o keep / (1 - p) with the stateless mask hash(seed_dev[0], salt, index) >= p (x, out: bf16 2-D)."""
    _chk_cuda(x, seed_dev, out)
    assert x.dim() == 2 and x.stride(1) == 1 and seed_dev.dtype == torch.int64
    out = torch.empty(x.shape, dtype=BF16, device=x.device) if out is None else out
    lib = _lib.load()
    _lib.check(lib.csm_lora_dropout(_p(x), _p(out), x.shape[0], x.shape[1], x.stride(0), out.stride(0), float(p),
                                    _p(seed_dev), int(salt), 1 if accumulate else 0, _st()), "lora_dropout")
    return out


# ----------------------------------------------------------------------------- helpers
def f32_to_bf16_(src_f32, dst_bf16, scale: float = 1.0, accumulate: bool = False):
    _chk_cuda(src_f32, dst_bf16)
    lib = _lib.load()
    _lib.check(lib.csm_f32_to_bf16(_p(src_f32), _p(dst_bf16), src_f32.numel(), scale, 1 if accumulate else 0, _st()),
               "f32_to_bf16")
    return dst_bf16


def add_bf16(a, b, out=None):
    _chk_cuda(a, b, out)
    out = torch.empty_like(a) if out is None else out
    lib = _lib.load()
    _lib.check(lib.csm_add_bf16(_p(a), _p(b), _p(out), a.numel(), _st()), "add_bf16")
    return out
