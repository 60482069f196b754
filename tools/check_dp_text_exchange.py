#!/usr/bin/env python
"""2+ GPUs under torchrun: the row-sparse text-embedding gradient exchange (dp.GradSynchronizer.exchange_text_rows)
against the dense path (local scatter + all-reduce(AVG)) on per-rank random frames.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dp_text_exchange.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "csm-train-pytorch_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from csm import ops  # noqa: E402
from csm.training import dp  # noqa: E402

rank, world, local = dp.init_distributed()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
Vt, D, B, S, C = 5000, 256, 2, 64, 32
g = torch.Generator().manual_seed(100 + rank)
tokens = torch.randint(0, 200, (B, S, C + 1), generator=g)
tokens[..., C] = torch.randint(0, Vt, (B, S), generator=g)
mask = torch.zeros(B, S, C + 1, dtype=torch.bool)
mask[:, : S // 4, C] = True                     # text frames
mask[:, S // 4:, :C] = True                     # audio frames
dh = torch.randn(B, S, D, generator=g).to(torch.bfloat16)
tokens, mask, dh = tokens.to(dev), mask.to(dev), dh.to(dev)
table = torch.nn.Parameter(torch.zeros(Vt, D, dtype=torch.bfloat16, device=dev))
other = torch.nn.Parameter(torch.zeros(8, 8, dtype=torch.bfloat16, device=dev))
sync = dp.GradSynchronizer([table, other], bucket_bytes=1 << 20, sparse_rows=table)
assert sync.sparse_param is table and all(p is not table for p in sync.params)
got = sync.exchange_text_rows(tokens, mask, dh, (Vt, D))
ref = torch.zeros(Vt, D, dtype=torch.bfloat16, device=dev)
ops.embed_gather_sum_bwd(tokens, mask, dh, None, ref, 0, Vt)
ref32 = ref.float()
dist.all_reduce(ref32, op=dist.ReduceOp.SUM)
ref32 /= world
err = (got.float() - ref32).abs().max().item()
scale = ref32.abs().max().item()
touched = int((ref32.abs().sum(dim=1) > 0).sum())
if rank == 0:
    print(f"text-grad exchange: world={world} rows touched={touched} max|diff|={err:.3e} (max|ref|={scale:.3e})")
assert err <= 2e-2 * scale + 1e-6, (err, scale)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
