#!/usr/bin/env python
"""Experiment behind the fp32 residual stream: stock torch bf16 (parameters / GEMM inputs in bf16) with the residual
stream of both transformer stacks kept in fp32, compared per tensor with the fp32 oracle.  If the q/k-projection
gradient cosines recover, the bf16 rounding of the residual stream is what limits them.
   python tools/parity_probe_residual.py c2|c3 [B]"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "csm-train-pytorch_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import test_parity_csm1b_gpu as T  # noqa: E402

BF = torch.bfloat16


def layer_fwd(self, x, *, mask=None, input_pos=None):            # x: fp32 residual stream
    n = self.sa_norm(x).to(BF)
    h = self.attn(n, n, mask=mask, input_pos=input_pos).float() + x
    return h + self.mlp(self.mlp_norm(h).to(BF)).float()


def stack_fwd(self, tokens, *, mask=None, input_pos=None):
    h = tokens.float()
    for layer in self.layers:
        h = layer(h, mask=mask, input_pos=input_pos)
    return self.norm(h).float()


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device("cuda:0")
    c = T.CASES[name]
    orc, prod, cfg = T._build_pair(dev, c["mode"], c["r"], c["targets"], c["max_seq"])
    del prod
    batch = T._batch(cfg, B, c["S"], seed=4321 + B)
    o_loss, o_per = T._oracle_step(orc, batch, dev)
    ref = {n: q.grad.detach().float().clone() for n, q in orc.named_parameters() if q.grad is not None}
    T._zero_grads(orc)
    req = {n: q.requires_grad for n, q in orc.named_parameters()}
    stock = orc.to(BF)
    for n, q in stock.named_parameters():
        q.requires_grad_(req[n])
    res = {}
    for tag in ("bf16 residual (stock)", "fp32 residual"):
        if tag == "fp32 residual":
            for stack in (stock.backbone, stock.decoder):
                stack.forward = types.MethodType(stack_fwd, stack)
                for layer in stack.layers:
                    layer.forward = types.MethodType(layer_fwd, layer)
        T._zero_grads(stock)
        s_loss, s_per = T._oracle_step(stock, batch, dev)
        res[tag] = {n: q.grad.detach().float().clone() for n, q in stock.named_parameters() if q.grad is not None}
        print(f"{tag}: loss {float(s_loss):.4f} (fp32 {float(o_loss):.4f})")
    rows = []
    for n in ref:
        a = ref[n].flatten()
        rows.append((float(F.cosine_similarity(a, res["bf16 residual (stock)"][n].flatten(), dim=0)),
                     float(F.cosine_similarity(a, res["fp32 residual"][n].flatten(), dim=0)), n))
    rows.sort()
    print(f"{'cos(stock bf16)':>16} {'cos(fp32 resid)':>16}  tensor")
    for a, b, n in rows[:20]:
        print(f"{a:16.6f} {b:16.6f}  {n}")
    print(f"min cosine: stock bf16 {min(r[0] for r in rows):.6f}, fp32 residual {min(r[1] for r in rows):.6f}; "
          f"below 0.999: {sum(r[0] < 0.999 for r in rows)} vs {sum(r[1] < 0.999 for r in rows)} of {len(rows)}")


if __name__ == "__main__":
    main()
