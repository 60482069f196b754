"""RoPE table for the CUDA kernel (csm_rope): fp32 [max_seq, head_dim/2, (cos, sin)].

Follows torchtune 0.4.0 ``Llama3ScaledRoPE`` as configured by the reference
(/root/reference/src/csm/models/model.py:13-25: rope_base=500_000, scale_factor=32): Llama-3 frequency scaling
with low_freq_factor=1, high_freq_factor=4, old_context_len=8192; the rotation itself (adjacent pairs, fp32) is in
csrc/pointwise.cu::rope_kernel.
"""
import math

import torch


def llama3_scaled_freqs(head_dim: int, base: float, scale_factor: float, low: float = 1.0, high: float = 4.0,
                        old_context_len: int = 8192) -> torch.Tensor:
    freqs = 1.0 / (base ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    out = []
    for f in freqs.tolist():
        wavelen = 2 * math.pi / f
        if wavelen < old_context_len / high:
            out.append(f)
        elif wavelen > old_context_len / low:
            out.append(f / scale_factor)
        else:
            smooth = (old_context_len / wavelen - low) / (high - low)
            out.append((1 - smooth) * f / scale_factor + smooth * f)
    return torch.tensor(out, dtype=freqs.dtype)


def build_rope_cache(head_dim: int, max_seq_len: int, base: float = 500_000.0,
                     scale_factor: float = 32.0) -> torch.Tensor:
    theta = llama3_scaled_freqs(head_dim, base, scale_factor)
    pos = torch.arange(max_seq_len, dtype=theta.dtype)
    ang = torch.einsum("i,j->ij", pos, theta).float()
    return torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).contiguous()
