"""Event-timed causal GQA attention backward (delta + dQ + dK/dV kernels) at the CSM-1B backbone shape, with the inverse
RoPE fused, as the training step calls it.   CSM_ATTN_BWD_OVERLAP=0/1 python tools/bench_attn_bwd.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "csm-train-pytorch_b200"))
import torch  # noqa: E402

from csm import ops  # noqa: E402
from csm.models.rope import build_rope_cache  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
out = {"overlap": os.environ.get("CSM_ATTN_BWD_OVERLAP", "0")}
flush = torch.zeros(512 << 20, dtype=torch.uint8, device=dev).view(torch.int32)
for B, S in ((2, 2048), (1, 4096), (2, 4096)):
    H, KV, hd = 32, 8, 64
    torch.manual_seed(0)
    qkv = torch.randn(B * S, (H + 2 * KV) * hd, device=dev).to(BF)
    q, k, v = qkv[:, :H * hd], qkv[:, H * hd:(H + KV) * hd], qkv[:, (H + KV) * hd:]
    do = torch.randn(B * S, H * hd, device=dev).to(BF)
    cache = build_rope_cache(hd, S).to(dev)
    o, lse = ops.attention_fwd(q, k, v, B, S, H, KV, hd)
    dqkv = torch.empty_like(qkv)
    fn = lambda: ops.attention_bwd(q, k, v, o, lse, do, B, S, H, KV, hd, dq=dqkv[:, :H * hd],  # noqa: E731
                                   dk=dqkv[:, H * hd:(H + KV) * hd], dv=dqkv[:, (H + KV) * hd:], rope_cache=cache)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.max(); flush.max()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    out[f"B{B}_S{S}_bwd_us"] = round(ts[len(ts) // 2], 1)
    out[f"B{B}_S{S}_checksum"] = float(dqkv.float().abs().sum())
print(json.dumps(out))
