"""CUDA-graph replay of a whole optimiser step (forward + backward + gradient exchange + clip + AdamW).

A step is ~700 kernel launches issued from Python through ctypes; once shapes are static the launch sequence is
captured once and replayed, so the step costs GPU time only ("CUDA streams and graphs instead of a tracing
compiler").  The first ``warmup`` calls run eagerly (they configure kernels and allocate optimiser state); the next
call captures and replays.  A batch whose shapes differ from the captured ones runs eagerly.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch


class GraphedStep:
    def __init__(self, step_fn: Callable[[Dict[str, torch.Tensor]], torch.Tensor], device, warmup: int = 3):
        self.step_fn = step_fn
        self.device = torch.device(device)
        self.warmup = warmup
        self.calls = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static: Optional[Dict[str, torch.Tensor]] = None
        self.loss: Optional[torch.Tensor] = None
        self.kernels_per_replay = 0

    def _signature(self, batch):
        return tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(batch.items()))

    def __call__(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        self.calls += 1
        if self.graph is None:
            if self.calls <= self.warmup:
                return self.step_fn({k: v.to(self.device, non_blocking=True) for k, v in batch.items()})
            self.static = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in batch.items()}
            self.sig = self._signature(batch)
            for k, v in batch.items():
                self.static[k].copy_(v, non_blocking=True)
            torch.cuda.synchronize(self.device)
            torch.cuda.empty_cache()      # the capture allocates the step's activations again in the graph's private pool:
            #                               hand the eager warm-up's cached blocks back first (long-context batches)
            from .. import _lib
            n0 = _lib.launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss = self.step_fn(self.static)
            self.kernels_per_replay = _lib.launch_count() - n0     # libcsm_b200 kernels recorded in the graph
            self.graph.replay()
            return self.loss.clone()
        if self._signature(batch) != self.sig:
            return self.step_fn({k: v.to(self.device, non_blocking=True) for k, v in batch.items()})
        for k, v in batch.items():
            self.static[k].copy_(v, non_blocking=True)      # H2D (pinned host batch) or D2D into the captured buffers
        self.graph.replay()
        return self.loss.clone()       # the captured loss buffer is overwritten by the next replay
