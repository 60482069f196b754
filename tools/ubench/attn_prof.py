"""Prints the wait-cycle counters of the profiling library (tools/ubench/attn_prof.sh) for one attention backward."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "csm-train-pytorch_b200"))
from csm import _lib, ops  # noqa: E402

B, S, H, KV, hd = 2, int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 32, 8, 64
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(B * S, H * hd, device="cuda", generator=g).bfloat16()
k = torch.randn(B * S, KV * hd, device="cuda", generator=g).bfloat16()
v = torch.randn(B * S, KV * hd, device="cuda", generator=g).bfloat16()
do = torch.randn(B * S, H * hd, device="cuda", generator=g).bfloat16()
o, lse = ops.attention_fwd(q, k, v, B, S, H, KV, hd)
for _ in range(3):
    ops.attention_bwd(q, k, v, o, lse, do, B, S, H, KV, hd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attention_bwd(q, k, v, o, lse, do, B, S, H, KV, hd)
e1.record()
torch.cuda.synchronize()
print(f"attention backward B={B} S={S}: {e0.elapsed_time(e1) * 100:.1f} us per call (delta + dq + dkdv, warm L2)")
lib = _lib.load()
buf = (C.c_longlong * 128)()
lib.csm_debug_attn_prof.restype = C.c_int
rc = lib.csm_debug_attn_prof(buf)
assert rc == 0, rc
p = list(buf)


def show(name, base, labels):
    tot = p[base]
    print(f"{name}: total {tot} cycles" + (f", {p[base + 4]} sub-blocks = {tot / max(p[base + 4], 1):.0f} cycles each" if "MMA" in name else ""))
    for i, lab in enumerate(labels):
        print(f"    {lab:<44} {p[base + 1 + i]:>9}  {100.0 * p[base + 1 + i] / max(tot, 1):5.1f} %")


show("dq   CTA 0  MMA warp", 0, ["wait K/V tile (TMA)", "wait S/dP buffer free (dQ MMA done)", "wait dS written (compute warps)"])
for w, base in (("2", 8), ("7", 16)):
    show(f"dq   CTA 0  compute warp {w}", base, ["wait S/dP ready (MMA)", "tcgen05.ld x2 + wait", "math + tcgen05.st + wait + arrive"])
show("dkdv CTA 0  MMA warp", 32, ["wait Q/dO tile (TMA)", "wait S/dP buffer free (dV/dK MMA done)", "wait P/dS written (compute warps)"])
for w, base in (("2", 40), ("7", 48)):
    show(f"dkdv CTA 0  compute warp {w}", base, ["wait S^T/dP^T ready (MMA)", "tcgen05.ld x2 + wait", "math + tcgen05.st x2 + wait + arrive", "lse/delta staging barrier"])

print("dq CTA 0 timeline (cycles since kernel entry): setup done %d, Q/dO in TMEM %d, first S/dP issued %d, first S/dP in registers %d, "
      "last dS written %d, dQ complete %d, exit %d" % tuple(p[64:71]))
