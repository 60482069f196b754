"""ctypes binding of libcsm_b200.so (include/csm_b200.h).

The library is the product: there is no CPU or eager-PyTorch fallback.  Loading fails loudly when the
shared object is missing, and every op raises ``RuntimeError`` when the C ABI returns a non-zero code.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CSM_B200_LIB", os.path.join(os.path.dirname(_HERE), "libcsm_b200.so"))

_lib = None
ABI_VERSION = 2     # include/csm_b200.h CSM_ABI_VERSION

_i32, _i64, _f32, _ptr, _sz = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); mirrors include/csm_b200.h one to one
SIGNATURES = {
    "csm_abi_version": (_i32, []),
    "csm_last_error": (C.c_char_p, []),
    "csm_device_supported": (_i32, []),
    "csm_launch_count": (_i64, []),
    "csm_embed_gather_sum_fwd": (_i32, [_ptr] * 8 + [_i64, _i32, _i64, _i64, _i32, _ptr]),
    "csm_embed_gather_sum_bwd": (_i32, [_ptr] * 5 + [_i64, _i32, _i64, _i64, _i32, _ptr]),
    "csm_embed_gather_sum_packed_fwd": (_i32, [_ptr] * 6 + [_i64, _i32, _i64, _i64, _i32, _ptr]),
    "csm_embed_gather_sum_packed_bwd": (_i32, [_ptr] * 5 + [_i64, _i32, _i64, _i64, _i32, _ptr]),
    "csm_decoder_input_fwd": (_i32, [_ptr] * 5 + [_i64, _i64, _i64, _i32, _i64, _i32, _ptr]),
    "csm_decoder_input_bwd": (_i32, [_ptr] * 5 + [_i64, _i64, _i64, _i32, _i64, _i32, _ptr]),
    "csm_rmsnorm_fwd": (_i32, [_ptr] * 4 + [_i64, _i32, _f32, _i32, _ptr]),
    "csm_rmsnorm_bwd": (_i32, [_ptr] * 7 + [_i64, _i32, _i32, _ptr]),
    "csm_rope": (_i32, [_ptr, _ptr, _i64, _i32, _i32, _i32, _i64, _i32, _ptr, _ptr]),
    "csm_gemm_bf16": (_i32, [_ptr] * 4 + [_i64] * 7 + [_i32] * 4 + [_f32, _ptr, _ptr, _i64, _i64, _i64, _i32, _ptr]),
    "csm_gemm_bf16_rope": (_i32, [_ptr] * 3 + [_i64] * 6 + [_ptr, _ptr, _i64, _i64, _i64, _ptr, _i32, _i32, _i32, _ptr, _ptr]),
    "csm_gemm_splitk_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "csm_gemm_bf16_splitk": (_i32, [_ptr] * 3 + [_i64] * 6 + [_i32, _i32, _f32, _i32, _ptr, _sz, _ptr]),
    "csm_gemm_swiglu_supported": (_i32, [_i64, _i64, _i64]),
    "csm_gemm_swiglu_fwd": (_i32, [_ptr] * 4 + [_i64] * 7 + [_ptr, _ptr, _i64, _i64, _i64, _ptr]),
    "csm_gemm_swiglu_bwd": (_i32, [_ptr] * 4 + [_i64] * 7 + [_ptr, _ptr, _i64, _i64, _i64, _ptr]),
    "csm_swiglu_fwd": (_i32, [_ptr] * 3 + [_i64] * 5 + [_ptr]),
    "csm_swiglu_bwd": (_i32, [_ptr] * 5 + [_i64] * 7 + [_ptr]),
    "csm_attn_causal_gqa_fwd": (_i32, [_ptr] * 5 + [_i32] * 5 + [_i64] * 4 + [_f32, _ptr]),
    "csm_set_attn_backend": (None, [_i32]),
    "csm_set_attn_fwd_variant": (None, [_i32]),
    "csm_set_gemm_cta_pair_mode": (None, [_i32]),
    "csm_set_gemm_dynamic_tiles": (None, [_i32]),
    "csm_set_reserved_sms": (None, [_i32]),
    "csm_gemm_experiments_compiled": (_i32, []),
    "csm_gemm_streamk_workspace_bytes": (_sz, []),
    "csm_gemm_set_streamk_workspace": (None, [_ptr, _sz]),
    "csm_set_gemm_streamk_mode": (None, [_i32]),
    "csm_set_gemm_narrow_tail_mode": (None, [_i32]),
    "csm_attn_bwd_workspace_bytes": (_sz, [_i32] * 5),
    "csm_attn_causal_gqa_bwd": (_i32, [_ptr] * 9 + [_i32] * 5 + [_i64] * 7 + [_f32, _ptr, _sz, _ptr]),
    "csm_attn_causal_gqa_bwd_rope": (_i32, [_ptr] * 9 + [_i32] * 5 + [_i64] * 7 + [_f32, _ptr, _ptr, _sz, _ptr]),
    "csm_linear_ce_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32]),
    "csm_linear_ce_fwd": (_i32, [_ptr] * 5 + [_i64] * 3 + [_i32] + [_i64] * 4 + [_i32, _i64, _i64, _ptr, _sz, _i32, _ptr]),
    "csm_linear_ce_bwd": (_i32, [_ptr] * 4 + [_f32, _ptr, _ptr, _ptr, _i32] + [_i64] * 3 + [_i32] + [_i64] * 4 +
                          [_i32] + [_i64] * 4 + [_ptr, _sz, _i32, _ptr]),
    "csm_adamw_clip_step": (_i32, [_ptr] * 7 + [_i32, _f32, _f32, _f32, _f32, _ptr, _ptr, _ptr]),
    "csm_adamw_clip_step_v2": (_i32, [_ptr] * 11 + [_i32, _f32, _f32, _f32, _f32, _i32, _ptr, _ptr, _ptr]),
    "csm_attn_varlen_fwd": (_i32, [_ptr] * 5 + [_i32] * 5 + [_i64] * 4 + [_f32, _ptr, _ptr]),
    "csm_attn_varlen_bwd": (_i32, [_ptr] * 9 + [_i32] * 5 + [_i64] * 7 + [_f32, _ptr, _ptr, _ptr, _ptr, _sz, _ptr]),
    "csm_attn_decode": (_i32, [_ptr] * 4 + [_i32] * 5 + [_i64] * 4 + [_f32, _ptr]),
    "csm_lora_mask_rows": (_i32, [_ptr, _i64, _i64, _i32, _ptr, _i32, _i32, _ptr]),
    "csm_lora_dropout": (_i32, [_ptr, _ptr, _i64, _i64, _i64, _i64, _f32, _ptr, _i64, _i32, _ptr]),
    "csm_set_pdl": (None, [_i32]),
    "csm_set_skinny_mode": (None, [_i32]),
    "csm_skinny_supported": (_i32, [_i32, _ptr, _ptr, _ptr] + [_i64] * 6 + [_i32]),
    "csm_skinny_rowdot": (_i32, [_ptr] * 3 + [_i64] * 6 + [_i32, _f32, _ptr]),
    "csm_skinny_coldot": (_i32, [_ptr] * 3 + [_i64] * 6 + [_i32, _f32, _ptr]),
    "csm_f32_to_bf16": (_i32, [_ptr, _ptr, _i64, _f32, _i32, _ptr]),
    "csm_add_bf16": (_i32, [_ptr, _ptr, _ptr, _i64, _ptr]),
}


def load():
    """Loads the shared library once; raises if it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"libcsm_b200.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or csm-train-pytorch_b200/build.sh). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch: fail loudly
        fn.restype, fn.argtypes = res, args
    if lib.csm_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libcsm_b200.so ABI version {lib.csm_abi_version()} != {ABI_VERSION}: rebuild it")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().csm_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().csm_launch_count())
