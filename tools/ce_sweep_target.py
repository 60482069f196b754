"""Launch target for an ncu launch list of the fused audio_head CE forward at N_sel = 64 / 128 / 232 / 256 (c5 sizes).
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --csv python tools/ce_sweep_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "csm-train-pytorch_b200"))
import torch  # noqa: E402

from csm import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
Dd, V, G = 1024, 2051, 31
torch.manual_seed(0)
PAD = int(os.environ.get("CE_PAD", "0"))          # extra elements per weight row (row pitch = (Dd + PAD) * 2 bytes)
head_t = (torch.randn(G, V, Dd + PAD, device=dev) * 0.05).to(BF)[:, :, :Dd]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for Ns in (64, 128, 232, 256):
    y = torch.randn(Ns, 32, Dd, device=dev).to(BF)
    codes = torch.randint(0, V, (Ns, 32), device=dev)
    for _ in range(3):
        flush.zero_()
        ops.linear_ce_fwd(y[:, 1:], head_t, codes[:, 1:], groups=G, tgt_row_stride=32, tgt_group_stride=1)
torch.cuda.synchronize()
print("done")
