import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "csm-train-pytorch_b200"))
import torch
from csm import ops
dev = torch.device("cuda:0"); BF = torch.bfloat16
g = torch.Generator().manual_seed(0)
for (N, R, IN) in [(256, 8, 256), (256, 8, 512), (256, 16, 256), (256, 24, 256)]:
    dts = torch.randn(N, R, generator=g).to(BF).to(dev)
    A = torch.randn(R, IN, generator=g).to(BF).to(dev)
    tmp = ops.gemm(dts, A, trans_b=True)
    ref = dts.float() @ A.float()
    print(N, R, IN, "gemm err", float((tmp.float() - ref).abs().max() / ref.abs().max()))
    seed = torch.ones(1, dtype=torch.int64, device=dev)
    dx = torch.randn(N, IN, generator=g).to(BF).to(dev)
    dx0 = dx.clone()
    keep = ops.lora_dropout(torch.ones(N, IN, dtype=BF, device=dev), 0.25, seed, 6).float()
    ops.lora_dropout(tmp, 0.25, seed, 6, out=dx, accumulate=True)
    want = dx0.float() + tmp.float() * (keep > 0) / 0.75
    print("  drop-acc err", float((dx.float() - want).abs().max() / want.abs().max()))
    xd = ops.lora_dropout(tmp, 0.25, seed, 6)
    print("  drop err", float((xd.float() - tmp.float() * (keep > 0) / 0.75).abs().max()))
