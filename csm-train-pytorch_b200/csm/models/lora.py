"""LoRA on the PyTorch/B200 ``Model``: semantics of /root/reference/src/csm/mlx/components/lora.py
(``LoRALinear`` :71-105, merge :140-153, targets :203-338, defaults :801-803, applied to backbone AND decoder
:805-827), re-homed from MLX.

An adapted projection keeps its frozen ``weight`` and gains ``lora_A`` [r, in] and ``lora_B`` [out, r] parameters
registered on the same ``nn.Linear`` module, so state-dict names are
``{backbone|decoder}.layers.{i}.attn.{q_proj,k_proj,v_proj,output_proj}.lora_{A,B}`` and
``...mlp.{w1,w3,w2}.lora_{A,B}`` (lora.py:115-119,215,272,291,310,329; the ``backbone.``/``decoder.`` prefix removes
the name collision of lora.py:832-844).  The low-rank term is computed inside the base GEMM's main loop
(csrc/gemm_tc.cu, extra K block) — never as separate unfused matmuls.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.nn as nn

TARGETS = {"q_proj": ("attn", "q_proj"), "k_proj": ("attn", "k_proj"), "v_proj": ("attn", "v_proj"),
           "o_proj": ("attn", "output_proj"), "gate_proj": ("mlp", "w1"), "up_proj": ("mlp", "w3"),
           "down_proj": ("mlp", "w2")}
DEFAULT_TARGETS = ["q_proj", "v_proj"]          # lora.py:801-803


def apply_lora(model: nn.Module, r: int = 8, alpha: float = 16.0, target_modules: Optional[Sequence[str]] = None,
               target_layers: Optional[Iterable[int]] = None, seed: Optional[int] = None,
               b_std: float = 0.0, num_adapters=1, target_decoder_layers: Optional[Iterable[int]] = None,
               dropout: float = 0.0, use_bias: bool = False) -> List[str]:
    """Freezes every base parameter and adds adapters (A ~ N(0, 1/sqrt(in)), B = 0 as lora.py:62-66; ``b_std`` > 0
    draws B ~ N(0, b_std) for gradient-parity tests, since B = 0 makes dA identically 0).

    ``num_adapters`` (int, or {"backbone": Kb, "decoder": Kd}) > 1 = multi-adapter batching, the GPU-native form of
    MultiSpeakerLoRATrainer (multi_speaker_lora.py:276-300): the K adapters of a projection sit side by side —
    ``lora_A`` [K*r, in] (rows k*r..(k+1)*r = adapter k), ``lora_B`` [out, K*r] — and ride in ONE base GEMM; each row
    of the batch keeps only its own adapter's block of t = x A^T (csm_lora_mask_rows).  ``target_layers`` filters the
    backbone layers (and the decoder's too unless ``target_decoder_layers`` is given), like
    target_backbone_layers / target_decoder_layers of the reference.
    ``dropout`` (lora.py:87-90): the low-rank path sees dropout(x, p) in training mode — a stateless hashed mask
    (csm_lora_dropout), re-derived in the backward.  ``use_bias`` (lora.py:66,101-102): a trainable ``lora_bias`` [out]
    per adapted projection, added unscaled; it rides in the base GEMM as one more tail column."""
    if not 0.0 <= dropout < 1.0:
        raise ValueError("lora dropout must be in [0, 1)")
    if r < 1 or r > 64:
        raise ValueError("lora_r must be in [1, 64]")
    if isinstance(num_adapters, int):
        num_adapters = {"backbone": num_adapters, "decoder": num_adapters}
    target_modules = list(target_modules or DEFAULT_TARGETS)
    for t in target_modules:
        if t not in TARGETS:
            raise ValueError(f"unknown LoRA target module {t!r}; choose from {sorted(TARGETS)}")
    for p in model.parameters():
        p.requires_grad_(False)
    g = torch.Generator(device="cpu")
    if seed is not None:
        g.manual_seed(seed)
    else:
        g.seed()
    layers_filter = {"backbone": None if target_layers is None else set(target_layers)}
    layers_filter["decoder"] = layers_filter["backbone"] if target_decoder_layers is None else set(target_decoder_layers)
    names = []
    for stack_name in ("backbone", "decoder"):
        stack = getattr(model, stack_name)
        K = int(num_adapters.get(stack_name, 1))
        if K < 1:
            raise ValueError("num_adapters must be >= 1")
        stack.lora_adapters = K
        stack.lora_dropout = float(dropout)
        for li, layer in enumerate(stack.layers):
            if layers_filter[stack_name] is not None and li not in layers_filter[stack_name]:
                continue
            for t in target_modules:
                parent_name, child = TARGETS[t]
                lin = getattr(getattr(layer, parent_name), child)
                w = lin.weight
                A = (torch.randn(K * r, lin.in_features, generator=g) / math.sqrt(lin.in_features))
                B = torch.randn(lin.out_features, K * r, generator=g) * b_std if b_std > 0 else \
                    torch.zeros(lin.out_features, K * r)
                lin.register_parameter("lora_A", nn.Parameter(A.to(device=w.device, dtype=w.dtype)))
                lin.register_parameter("lora_B", nn.Parameter(B.to(device=w.device, dtype=w.dtype)))
                lin.lora_scaling = alpha / r           # lora.py:52-53
                lin.lora_r, lin.lora_alpha, lin.lora_adapters = r, alpha, K
                if use_bias:
                    if K > 1:
                        raise ValueError("lora_use_bias is not supported with several adapters per projection")
                    lin.register_parameter("lora_bias", nn.Parameter(torch.zeros(lin.out_features, device=w.device,
                                                                                 dtype=w.dtype)))
                names.append(f"{stack_name}.layers.{li}.{parent_name}.{child}")
    return names


def lora_state_dict(model: nn.Module) -> Dict[str, torch.Tensor]:
    return {n: p.detach() for n, p in model.named_parameters() if n.endswith(("lora_A", "lora_B", "lora_bias"))}


def adapter_slices(mod: nn.Module, k: int):
    """(A_k [r, in], B_k [out, r]) views of adapter k of an adapted projection (B_k is row-strided for K > 1)."""
    r = mod.lora_r
    return mod.lora_A[k * r:(k + 1) * r], mod.lora_B[:, k * r:(k + 1) * r]


def merge_lora(model: nn.Module) -> None:
    """W0 + (alpha/r) B A, in place (lora.py:140-153), computed by the GEMM kernel; adapters are then zeroed (B=0).
    Single-adapter models only (with several adapters there is no one merged weight)."""
    from .. import ops
    for mod in model.modules():
        if isinstance(mod, nn.Linear) and hasattr(mod, "lora_A"):
            if getattr(mod, "lora_adapters", 1) != 1:
                raise RuntimeError("merge_lora: the model holds several adapters per projection; export one speaker "
                                   "(MultiSpeakerLoRATrainer.save_all_models) and merge that")
            with torch.no_grad():
                # W[out,in] += s * B[out,r] @ A[r,in]  ==  gemm(a=B [M=out,K=r], b=A stored [K=r, N=in] -> trans_b)
                ops.gemm(mod.lora_B.data, mod.lora_A.data, trans_b=True, out=mod.weight.data, accumulate=True,
                         alpha=float(mod.lora_scaling))
                mod.lora_B.zero_()
