"""Data-parallel gradient exchange: one process per GPU, full replica each, one all-reduce(sum) of the
trainable gradients per optimiser step over NCCL/NVLink, then x 1/world.  The reference has no multi-GPU code.

  * LoRA (<= ~20 M trainable params): one flat buffer, one latency-bound all-reduce after backward.
  * full fine-tune: reverse-order buckets launched from post-accumulate-grad hooks on a side stream so the exchange
    overlaps the remaining backward; the local sum of squares for global-norm clipping rides in the same pass.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None) -> tuple:
    """(rank, world, local_rank) from the torchrun environment; no-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


class GradSynchronizer:
    """Averages gradients of `params` across ranks.

    bucket_bytes=None -> a single flat bucket reduced in ``finish()`` (LoRA);
    otherwise reverse-order buckets reduced asynchronously as soon as all their gradients exist (full FT).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: Optional[int] = None,
                 group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.bucket_bytes = bucket_bytes
        self._buckets: List[List[torch.nn.Parameter]] = []
        self._pending = {}
        self._works = []
        self._hooks = []
        self._stream = None
        if self.world > 1 and bucket_bytes:
            cur, size = [], 0
            for p in reversed(self.params):                 # backward produces gradients roughly in reverse order
                cur.append(p)
                size += p.numel() * p.element_size()
                if size >= bucket_bytes:
                    self._buckets.append(cur)
                    cur, size = [], 0
            if cur:
                self._buckets.append(cur)
            self._bucket_of = {id(p): bi for bi, b in enumerate(self._buckets) for p in b}
            if self.params and self.params[0].is_cuda:
                self._stream = torch.cuda.Stream(priority=-1)
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
            self._reset()

    def _reset(self):
        self._pending = {bi: len(b) for bi, b in enumerate(self._buckets)}
        self._works = []

    def _reduce_bucket(self, bucket):
        grads = [p.grad for p in bucket if p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._works.append((work, flat, grads))

    def _on_grad(self, p):
        bi = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            if self._stream is not None:
                self._stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self._stream):
                    self._reduce_bucket(self._buckets[bi])
            else:
                self._reduce_bucket(self._buckets[bi])

    def finish(self) -> None:
        """Call after backward, before clipping / optimizer.step()."""
        if self.world == 1:
            return
        inv = 1.0 / self.world
        if not self.bucket_bytes:
            grads = [p.grad for p in self.params if p.grad is not None]
            if not grads:
                return
            flat = torch.cat([g.reshape(-1).float() for g in grads])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.mul_(inv)
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
            return
        for bi, left in self._pending.items():              # parameters that received no gradient this step
            if left > 0:
                self._reduce_bucket(self._buckets[bi])
        for work, flat, grads in self._works:
            work.wait()
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g)).mul_(inv)
                off += n
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        self._reset()
