#!/usr/bin/env python
"""Per-tensor gradient agreement with the fp32 oracle at CSM-1B dimensions: this repo's kernels AND stock PyTorch in
bf16 on the same GPU (the arithmetic the reference would run), side by side.  Diagnostic for tests/test_parity_csm1b_gpu.py.
   python tools/parity_probe.py c2|c3|c4 [B]"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "csm-train-pytorch_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import test_parity_csm1b_gpu as T  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device("cuda:0")
    c = T.CASES[name]
    orc, prod, cfg = T._build_pair(dev, c["mode"], c["r"], c["targets"], c["max_seq"])
    batch = T._batch(cfg, B, c["S"], seed=4321 + B)
    o_loss, o_per = T._oracle_step(orc, batch, dev)
    p_loss, p_per, _ = T._product_step(prod, batch, dev, graph=False)
    ref = {n: q.grad.detach().float().clone() for n, q in orc.named_parameters() if q.grad is not None}
    mine = {n: q.grad.detach().float() for n, q in prod.named_parameters() if q.grad is not None}
    T._zero_grads(orc)
    req = {n: q.requires_grad for n, q in orc.named_parameters()}
    stock = orc.to(torch.bfloat16)                       # same module, bf16 parameters / activations: stock torch bf16
    for n, q in stock.named_parameters():
        q.requires_grad_(req[n])
    s_loss, s_per = T._oracle_step(stock, batch, dev)
    theirs = {n: q.grad.detach().float() for n, q in stock.named_parameters() if q.grad is not None}
    rows = []
    for n in ref:
        a = ref[n].flatten()
        cm = float(F.cosine_similarity(a, mine[n].flatten(), dim=0))
        cs = float(F.cosine_similarity(a, theirs[n].flatten(), dim=0))
        rows.append((cm, cs, n, float(a.norm()), float(mine[n].norm()), float(theirs[n].norm())))
    rows.sort()
    print(f"{name} B={B}: loss ours {float(p_loss):.4f} stock-bf16 {float(s_loss):.4f} fp32 {float(o_loss):.4f}")
    print(f"per-codebook max rel err: ours {float(((p_per - o_per).abs() / o_per).max()):.2e}  "
          f"stock-bf16 {float(((s_per - o_per).abs() / o_per).max()):.2e}")
    print(f"{'cos(ours,fp32)':>15} {'cos(stock,fp32)':>16}  |g| fp32 / ours / stock   tensor")
    for cm, cs, n, na, nm, ns in rows[:25]:
        print(f"{cm:15.6f} {cs:16.6f}  {na:.3e} {nm:.3e} {ns:.3e}  {n}")
    below_m = sum(1 for r in rows if r[0] < 0.999)
    below_s = sum(1 for r in rows if r[1] < 0.999)
    print(f"tensors below 0.999: ours {below_m} / {len(rows)}, stock bf16 {below_s} / {len(rows)}")
    print(f"min cosine: ours {rows[0][0]:.6f}, stock bf16 {min(r[1] for r in rows):.6f}")


if __name__ == "__main__":
    main()
