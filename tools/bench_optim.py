#!/usr/bin/env python
"""Global-norm clip + AdamW kernels (csrc/optim.cu) on a CSM-1B-sized parameter set (1.55 B bf16 elements in tensors
of the model's shapes), CUDA events; prints achieved GB/s of the 16 B/param the two passes move, next to torch's
clip_grad_norm_ + fused AdamW on the same tensors.   python tools/bench_optim.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "csm-train-pytorch_b200"))
import torch  # noqa: E402

from csm.training.optim import FusedClipAdamW  # noqa: E402

dev = torch.device("cuda:0")
shapes = [(128256, 2048), (65632, 2048), (31, 1024, 2051), (2051, 2048), (1024, 2048)]
for _ in range(16):
    shapes += [(3072, 2048), (2048, 2048), (16384, 2048), (2048, 8192), (2048,), (2048,)]
for _ in range(4):
    shapes += [(1536, 1024), (1024, 1024), (16384, 1024), (1024, 8192), (1024,), (1024,)]
n = sum(torch.Size(s).numel() for s in shapes)


def make():
    return [torch.nn.Parameter(torch.randn(s, device=dev, dtype=torch.bfloat16) * 0.02) for s in shapes]


def time_steps(step, iters=5):
    for _ in range(2):
        step()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


ps = make()
for p in ps:
    p.grad = torch.randn_like(p)
ours = FusedClipAdamW(ps, lr=1e-4, weight_decay=0.01)
ms_ours = time_steps(lambda: ours.step(max_grad_norm=1.0))
del ours
ref = torch.optim.AdamW(ps, lr=1e-4, weight_decay=0.01, fused=True, capturable=True)


def ref_step():
    torch.nn.utils.clip_grad_norm_(ps, 1.0)
    ref.step()


ms_ref = time_steps(ref_step)
out = {"params": n, "ours_ms": ms_ours, "ours_GBs": 16.0 * n / ms_ours / 1e6, "torch_ms": ms_ref,
       "torch_GBs_equiv": 16.0 * n / ms_ref / 1e6}
print(f"{n/1e9:.2f} B bf16 params: clip + AdamW  ours {ms_ours:.2f} ms ({out['ours_GBs']:.0f} GB/s of 16 B/param)   "
      f"torch clip_grad_norm_ + fused AdamW {ms_ref:.2f} ms")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "kernels_optim.json"), "w"), indent=1)
