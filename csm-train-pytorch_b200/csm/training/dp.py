"""Data-parallel gradient exchange: one process per GPU, full replica each, one all-reduce(mean) of the
trainable gradients per optimiser step over NCCL/NVLink.  The reference has no multi-GPU code.

  * LoRA (<= ~20 M trainable params): one flat fp32 buffer, one latency-bound all-reduce after backward.
  * full fine-tune: gradients live in flat bf16 **bucket buffers** laid out in backward order; ``param.grad`` is a view
    into its bucket.  The transformer stack's hand-written backward delivers each layer's gradients as soon as the
    layer is done (``deliver``), every other parameter arrives through a post-accumulate-grad hook; a bucket is
    all-reduced in place on a high-priority side stream the moment its last gradient lands, so the exchange overlaps
    the remaining backward (and is captured in the step's CUDA graph as a fork/join).
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional

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
    otherwise flat bucket buffers (``param.grad`` = view) reduced asynchronously as they fill (full fine-tune).
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: Optional[int] = None, group=None,
                 force_buckets: bool = False, sparse_rows: Optional[torch.nn.Parameter] = None,
                 bucket_order: Optional[List[torch.nn.Parameter]] = None):
        """`sparse_rows`: a parameter (the text-embedding table) whose gradient is exchanged as gathered rows by
        ``exchange_text_rows`` from inside the embedding backward; it is kept out of the all-reduce buckets.
        `bucket_order`: the parameters in the order the backward produces their gradients (bucket layout order);
        default: reverse parameter order.  Projections that share a fused GEMM (q|k|v, gate|up) should be adjacent and
        in GEMM order, so that one wgrad GEMM writes all of them straight into the bucket (``packed_view``)."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.bucketed = bool(bucket_bytes) and (self.world > 1 or force_buckets)
        self.sparse_param = None
        if self.bucketed and self.world > 1 and sparse_rows is not None and sparse_rows.requires_grad and \
                sparse_rows.is_cuda and os.environ.get("CSM_DP_DENSE_TEXT_GRAD", "0") != "1":
            self.sparse_param = sparse_rows
            self.params = [p for p in self.params if p is not sparse_rows]
        self.text_capacity_seq = 2048        # frames per sample the text-row exchange is sized for (max_seq_len)
        self.text_capacity_frames = None     # fixed by the first exchange
        self.trace = None                    # a list while bench.py times the exchange: ("bucket", i, bytes, e0, e1) /
        #                                      ("finish", e0, e1) CUDA-event records of one step
        self.accumulating = False            # True on all but the last micro-batch of an accumulation window
        self._hooks = []
        if not self.bucketed:
            return
        backend = dist.get_backend(group) if dist.is_initialized() else "none"
        self._avg = backend == "nccl"        # NCCL reduces with AVG in one pass; gloo needs SUM + scale
        self._buckets: List[Dict] = []
        self._slot: Dict[int, tuple] = {}    # id(param) -> (bucket index, grad view)
        cur, size = [], 0
        if bucket_order is not None:
            mine = {id(p) for p in self.params}
            order = [p for p in bucket_order if id(p) in mine]
            listed = {id(p) for p in order}
            order += [p for p in reversed(self.params) if id(p) not in listed]
        else:
            order = list(reversed(self.params))  # backward produces gradients roughly in reverse parameter order
        self._offset: Dict[int, tuple] = {}  # id(param) -> (bucket index, element offset)

        def close(ps):
            n = sum(p.numel() for p in ps)
            buf = torch.zeros(n, dtype=ps[0].dtype, device=ps[0].device)
            bi, off = len(self._buckets), 0
            for p in ps:
                self._slot[id(p)] = (bi, buf[off:off + p.numel()].view(p.shape))
                self._offset[id(p)] = (bi, off)
                off += p.numel()
            self._buckets.append({"buf": buf, "n": len(ps), "params": ps})
        for p in order:
            if cur and (cur[0].dtype != p.dtype or cur[0].device != p.device):
                close(cur)
                cur, size = [], 0
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                close(cur)
                cur, size = [], 0
        if cur:
            close(cur)
        self._stream = torch.cuda.Stream(priority=-1) if self.params[0].is_cuda else None
        if self.world > 1 and self.params[0].is_cuda and os.environ.get("CSM_DP_STATIC_TILES", "0") != "1":
            # the bucket all-reduces run beside the backward and hold SMs: a persistent GEMM with a static tile
            # assignment then waits for CTAs that could not start (measured at 8 GPUs: +6.6 ms per 41.8 ms step while
            # 10.9 ms of NCCL kernels overlap the backward); with tiles drawn from a counter late CTAs find less work
            from .. import ops
            ops.set_gemm_dynamic_tiles(1)
        if self.world > 1 and self.params[0].is_cuda:
            # the all-reduce CTAs run beside the backward: keep SMs free for them so the persistent GEMM grids never
            # wait for an SM that NCCL holds (CSM_DP_RESERVED_SMS, default 0 = off; pair with NCCL_MAX_CTAS)
            n = int(os.environ.get("CSM_DP_RESERVED_SMS", "0"))
            if n > 0:
                from .. import _lib
                _lib.load().csm_set_reserved_sms(n)
        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_autograd_grad))
        self._reset()

    # ------------------------------------------------------------------ bucketed mode
    def _reset(self):
        self._pending = [b["n"] for b in self._buckets]
        self._arrived = set()
        self._works = []

    def _launch(self, bi):
        if self.world == 1:
            return
        buf = self._buckets[bi]["buf"]
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                if self.trace is not None:
                    # timed form: the side stream itself waits for the collective, so two events on it bracket the
                    # all-reduce (plus its queueing behind earlier buckets)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    dist.all_reduce(buf, op=op, group=self.group)
                    e1.record()
                    self.trace.append(("bucket", bi, buf.numel() * buf.element_size(), e0, e1))
                else:
                    self._works.append(dist.all_reduce(buf, op=op, group=self.group, async_op=True))
        else:
            self._works.append(dist.all_reduce(buf, op=op, group=self.group, async_op=True))

    def _store(self, p, g) -> None:
        bi, view = self._slot[id(p)]
        if p.grad is view:                   # accumulation window: gradient of an earlier micro-batch is in the view
            view.add_(g)
        else:
            view.copy_(g)
            p.grad = view
        self._mark(p, bi)

    def _mark(self, p, bi) -> None:
        if self.accumulating or id(p) in self._arrived:
            return
        self._arrived.add(id(p))
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    # ---- gradients written straight into the bucket by the producing GEMM (no copy)
    def grad_view(self, p: torch.nn.Parameter) -> Optional[torch.Tensor]:
        """The bucket view that IS `p`'s gradient storage (None when `p` is not bucketed): a wgrad GEMM passes it as its
        output (accumulate = ``has_grad(p)``) and then calls ``mark_written(p)``."""
        if not self.bucketed or id(p) not in self._slot:
            return None
        return self._slot[id(p)][1]

    def packed_view(self, ps: List[torch.nn.Parameter]) -> Optional[torch.Tensor]:
        """One [sum(rows), cols] view over the adjacent bucket slots of several 2-D parameters with the same number of
        columns (q|k|v, gate|up), or None if they are not adjacent in one bucket in this order."""
        if not self.bucketed or any(id(p) not in self._offset for p in ps):
            return None
        bi, off0 = self._offset[id(ps[0])]
        off, cols = off0, ps[0].shape[1]
        for p in ps:
            b, o = self._offset[id(p)]
            if b != bi or o != off or p.dim() != 2 or p.shape[1] != cols:
                return None
            off += p.numel()
        return self._buckets[bi]["buf"][off0:off].view(-1, cols)

    def has_grad(self, p: torch.nn.Parameter) -> bool:
        """True inside an accumulation window when an earlier micro-batch's gradient already sits in the view."""
        return p.grad is not None and p.grad is self._slot[id(p)][1]

    def mark_written(self, p: torch.nn.Parameter) -> None:
        bi, view = self._slot[id(p)]
        p.grad = view
        self._mark(p, bi)

    def deliver(self, p: torch.nn.Parameter, g: torch.Tensor) -> bool:
        """Called by a hand-written backward with a FINAL gradient for `p`.  Returns True if the synchroniser took
        ownership (the caller must then return None for `p` to autograd)."""
        if not self.bucketed or id(p) not in self._slot:
            return False
        self._store(p, g)
        return True

    def _on_autograd_grad(self, p):
        bi, view = self._slot[id(p)]
        if p.grad is view:                   # autograd accumulated in place into the bucket view
            self._mark(p, bi)
            return
        g = p.grad
        p.grad = None
        self._store(p, g)

    # ------------------------------------------------------------------ row-sparse text-embedding gradient
    def exchange_text_rows(self, tokens, mask, dh, table_shape) -> torch.Tensor:
        """Called by EmbedGatherSumFn.backward: returns the dense, rank-averaged text-embedding gradient.
        Every rank contributes its frames (tokens [B*S,C+1], mask [B*S,C+1], dh [B*S,D]) padded to a FIXED capacity
        (``text_capacity_frames``); after the all-gathers each rank runs the embedding scatter kernel over all ranks'
        frames (text column only).  The collective sequence and every message size are therefore the same on every rank
        and in every step — whether a rank replays its CUDA graph or runs eagerly, and however ragged the batches are
        (a data-dependent shape exchange in the eager path only would desynchronise NCCL)."""
        from .. import ops
        W = self.world
        dt = torch.zeros(table_shape, dtype=dh.dtype, device=dh.device)   # fresh: autograd may adopt it as .grad
        msk = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8).contiguous()
        tokens, msk, dh = self.pad_to_capacity(tokens, msk, dh)
        tok_all = torch.empty((W,) + tuple(tokens.shape), dtype=tokens.dtype, device=tokens.device)
        msk_all = torch.empty((W,) + tuple(msk.shape), dtype=torch.uint8, device=tokens.device)
        dh_all = torch.empty((W,) + tuple(dh.shape), dtype=dh.dtype, device=dh.device)
        dist.all_gather_into_tensor(tok_all, tokens, group=self.group)
        dist.all_gather_into_tensor(msk_all, msk, group=self.group)
        dist.all_gather_into_tensor(dh_all, dh, group=self.group)
        dh_all.mul_(1.0 / W)                                     # mean over the ranks, like the bucket all-reduce(AVG)
        F, Wc = tokens.shape
        ops.embed_gather_sum_bwd(tok_all.view(W, F, Wc), msk_all.view(W, F, Wc), dh_all.view(W, F, -1),
                                 None, dt, 0, table_shape[0])
        return dt

    def pad_to_capacity(self, tokens, msk, dh):
        """[B,S,..] -> [capacity, ..] frame lists, padded with masked-out frames (mask 0, dh 0: they scatter nothing).
        The capacity is fixed by the first exchange (frames of that batch rounded up to whole sequences of
        ``text_capacity_seq`` frames, default 2048 = the model's max_seq_len), identically on every rank as long as
        the ranks use the same batch size — no shape collective, no host read."""
        B, S = tokens.shape[0], tokens.shape[1]
        frames = B * S
        if self.text_capacity_frames is None:
            seq = max(int(self.text_capacity_seq), 1)
            self.text_capacity_frames = B * max(seq, -(-S // seq) * seq)
        cap = self.text_capacity_frames
        if frames > cap:
            raise RuntimeError(
                f"data-parallel text-embedding exchange: batch of {B} x {S} = {frames} frames exceeds the fixed capacity "
                f"of {cap} frames set by the first step (every rank must use the same batch size; raise "
                f"GradSynchronizer.text_capacity_seq for sequences longer than {self.text_capacity_seq})")
        tokens, msk, dh = (t.reshape((frames,) + tuple(t.shape[2:])) for t in (tokens, msk, dh))
        if frames == cap:
            return tokens.contiguous(), msk.contiguous(), dh.contiguous()

        def grow(t):
            out = t.new_zeros((cap,) + tuple(t.shape[1:]))
            out[:frames] = t
            return out
        return grow(tokens), grow(msk), grow(dh)

    # ------------------------------------------------------------------ both modes
    def finish(self) -> None:
        """Call after backward, before clipping / optimizer.step()."""
        if self.trace is not None and (self.world > 1 or self.bucketed):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            tr, self.trace = self.trace, None          # (the nested call below must not re-enter this branch)
            try:
                self._finish()
            finally:
                self.trace = tr
            e1.record()
            self.trace.append(("finish", e0, e1))
            return
        self._finish()

    def _finish(self) -> None:
        if self.world == 1 and not self.bucketed:
            return
        if not self.bucketed:
            # one flat fp32 buffer over ALL trainable parameters (fixed layout, kept across steps), filled and drained
            # by multi-tensor copies.  A parameter without a gradient on this rank (e.g. the decoder adapters when the
            # batch selected no decoder frame) contributes zeros: every rank issues the same all-reduce with the same
            # count in every step.
            if not self.params:
                return
            flat = getattr(self, "_flat", None)
            if flat is None or flat.device != self.params[0].device:
                n = sum(p.numel() for p in self.params)
                flat = self._flat = torch.empty(n, dtype=torch.float32, device=self.params[0].device)
                self._flat_views, off = [], 0
                for p in self.params:
                    self._flat_views.append(flat[off:off + p.numel()].view(p.shape))
                    off += p.numel()
            have = [(v, p.grad) for v, p in zip(self._flat_views, self.params) if p.grad is not None]
            if len(have) != len(self.params):
                flat.zero_()
            if have:
                torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
            if dist.get_backend(self.group) == "nccl":
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                flat.mul_(1.0 / self.world)
            if have:
                torch._foreach_copy_([g for _, g in have], [v for v, _ in have])
            for v, p in zip(self._flat_views, self.params):
                if p.grad is None:
                    p.grad = v.to(p.dtype)
            return
        if self.accumulating:
            return
        for bi, left in enumerate(self._pending):      # parameters that received no gradient this step count as zero
            if left > 0:
                for p in self._buckets[bi]["params"]:
                    if id(p) not in self._arrived:
                        _, view = self._slot[id(p)]
                        view.zero_()
                        p.grad = view
                self._launch(bi)
        for w in self._works:
            w.wait()
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        if self.world > 1 and not self._avg:
            for b in self._buckets:
                b["buf"].mul_(1.0 / self.world)
        self._reset()
