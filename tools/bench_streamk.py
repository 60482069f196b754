#!/usr/bin/env python
"""A/B of the stream-K schedule of the CTA-pair GEMM (CUDA events, L2 flushed): python tools/bench_streamk.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "csm-train-pytorch_b200"))
import torch  # noqa: E402

from csm import ops  # noqa: E402
from tools.bench_kernels import timeit  # noqa: E402

dev = torch.device("cuda:0")
for M, N, K, tb in [(4096, 2048, 8192, 0), (4096, 2048, 16384, 1), (4096, 2048, 3072, 1), (4096, 2048, 2048, 0),
                    (7424, 1024, 8192, 0)]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    b = torch.randn((K, N) if tb else (N, K), device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    res = {}
    for mode in (0, 1, 0, 1):
        ops.set_gemm_streamk_mode(mode)
        res.setdefault(mode, []).append(timeit(lambda: ops.gemm(a, b, trans_b=bool(tb), out=out), iters=20))
    ops.set_gemm_streamk_mode(0)
    fl = 2.0 * M * N * K
    print(f"M={M} N={N} K={K} tb={tb}: whole tiles {min(res[0])*1e3:.1f} us ({fl/min(res[0])/1e9:.0f} TF/s)   "
          f"stream-K {min(res[1])*1e3:.1f} us ({fl/min(res[1])/1e9:.0f} TF/s)", flush=True)
