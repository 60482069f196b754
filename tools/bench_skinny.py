"""Per-kernel timings of the LoRA tall-skinny products (csrc/skinny.cu vs the tcgen05 split-reduction path) and of the
RMSNorm backward with / without the scale gradient.  CUDA events, 256 MB L2 flush before every timed launch, and a
second column with the operands left in L2 (what the step sees: the previous kernel just wrote them).
    python tools/bench_skinny.py > gpurun_out/skinny.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "csm-train-pytorch_b200"))
from csm import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


REPS = 10


def _capture(body):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    return g


def _graph_ms(g, iters=7):
    g.replay()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def _flush_only():
    for _ in range(REPS):
        flush.zero_()


_flush_ms = None


def per_op_us(fn):
    """(cold, warm) microseconds per call: a graph of REPS x [256 MB flush, op] minus a graph of REPS x [flush]; and a
    graph of REPS x [op] back to back (operands L2-resident).  Graph replays take the host launch path out of the number."""
    global _flush_ms
    if _flush_ms is None:
        _flush_ms = _graph_ms(_capture(_flush_only))

    def cold():
        for _ in range(REPS):
            flush.zero_()
            fn()

    def warm():
        for _ in range(REPS):
            fn()
    c = (_graph_ms(_capture(cold)) - _flush_ms) / REPS * 1e3
    w = _graph_ms(_capture(warm)) / REPS * 1e3
    return round(c, 2), round(w, 2)


g = torch.Generator(device=dev).manual_seed(0)
N, D, R = 4096, 2048, 16
x = torch.randn(N, D, device=dev, generator=g).to(BF)
dy = torch.randn(N, 3072, device=dev, generator=g).to(BF)
A = torch.randn(R, D, device=dev, generator=g).to(BF)
Bm = torch.randn(3072, R, device=dev, generator=g).to(BF)
t = torch.randn(N, R, device=dev, generator=g).to(BF)
cases = {
    "t = x A^T            [4096x2048].[16x2048]^T": lambda: ops.gemm(x, A, alpha=2.0),
    "dts = dy B           [4096x3072].[3072x16]": lambda: ops.gemm(dy, Bm, trans_b=True, alpha=2.0),
    "dB = dy^T t          [3072x16]": lambda: ops.gemm(dy, t, trans_a=True, trans_b=True),
    "dA = dts^T x         [16x2048]": lambda: ops.gemm(t, x, trans_a=True, trans_b=True),
}
out = {"skinny": {}, "rmsnorm_bwd": {}}
for name, fn in cases.items():
    row = {}
    for label, on in (("skinny", True), ("tcgen05_splitk", False)):
        ops.SKINNY_ENABLED = on
        row[label + "_cold_us"], row[label + "_l2_us"] = per_op_us(fn)
    ops.SKINNY_ENABLED = True
    out["skinny"][name] = row

xf = torch.randn(N, D, device=dev, generator=g)
scale = torch.ones(D, device=dev, dtype=BF)
dyn = torch.randn(N, D, device=dev, generator=g).to(BF)
dres = torch.randn(N, D, device=dev, generator=g).to(BF)
_, rstd = ops.rmsnorm(xf, scale, 1e-5)
ds = torch.zeros(D, device=dev)
for label, dsarg in (("no_dscale", None), ("with_dscale", ds)):
    c, w = per_op_us(lambda: ops.rmsnorm_bwd(dyn, xf, scale, rstd, dres, dsarg))
    out["rmsnorm_bwd"][label] = {"cold_us": c, "l2_us": w, "bytes": N * D * (4 + 2 + 2 + 2)}
c, w = per_op_us(lambda: ops.rmsnorm(xf, scale, 1e-5))
out["rmsnorm_fwd"] = {"cold_us": c, "l2_us": w, "bytes": N * D * (4 + 2)}
out["timing"] = "graph of 10 x [256 MB L2 flush, op] minus graph of 10 x [flush] (cold) / graph of 10 x [op] (operands in L2)"
print(json.dumps(out, indent=1))
