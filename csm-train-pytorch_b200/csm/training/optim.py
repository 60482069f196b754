"""Global-norm clipping + AdamW as two multi-tensor kernel passes (csrc/optim.cu).

Drop-in for ``clip_grad_norm_(params, max_norm); torch.optim.AdamW(...).step()`` of the reference trainer
(trainer.py:123-173, 269-278): same update rule (decoupled weight decay, bias correction, eps outside the square root),
same per-group learning rates, state tensors ``exp_avg`` / ``exp_avg_sq`` and a ``step`` counter, so ``state_dict()``
round-trips through the reference checkpoint format.  The clip coefficient is derived on the device from the
squared-norm accumulator, so the step (including the norm) replays inside a CUDA graph.

Precision (reference: fp32 parameters, fp32 AdamW, trainer.py:107,166-173): the bf16 parameters the kernels read are
the rounded image of an **fp32 master copy** kept in the optimiser state (``state[p]["master"]``) and the moments are
fp32 — at the reference learning rates (1e-5 x 0.1 for the backbone) an update is ~1/60 of a bf16 half-ulp and would
otherwise be rounded away.  ``master_weights=False`` / ``state_dtype=torch.bfloat16`` select the narrower round-1 modes.
Parameters may be row-strided 2-D views (LoRA B blocks of a block-diagonal operand, csm/autograd.py).
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Optional

import torch

from .. import _lib


class FusedClipAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_grad_norm: Optional[float] = None, master_weights: bool = True,
                 state_dtype: torch.dtype = torch.float32):
        if state_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("FusedClipAdamW: state_dtype must be torch.float32 or torch.bfloat16")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.master_weights = bool(master_weights)
        self.state_dtype = state_dtype
        b = {tuple(g["betas"]) for g in self.param_groups}
        e = {float(g["eps"]) for g in self.param_groups}
        if len(b) != 1 or len(e) != 1:
            raise ValueError("FusedClipAdamW: betas and eps must be the same in every parameter group")
        self.max_grad_norm = max_grad_norm
        self._dev_scalars = None          # [step, squared gradient norm] fp32, on the parameters' device
        self._tables = None               # ctypes arrays of the last step (rebuilt when a pointer changes)
        self._table_key = None

    # ------------------------------------------------------------------ state
    def _scalars(self, device):
        if self._dev_scalars is None or self._dev_scalars.device != device:
            self._dev_scalars = torch.zeros(2, dtype=torch.float32, device=device)
        return self._dev_scalars

    def _init_state(self, p):
        st = self.state[p]
        if "exp_avg" not in st:
            st["exp_avg"] = torch.zeros(p.shape, dtype=self.state_dtype, device=p.device)
            st["exp_avg_sq"] = torch.zeros(p.shape, dtype=self.state_dtype, device=p.device)
        if self.master_weights and "master" not in st:
            st["master"] = p.detach().to(torch.float32).contiguous()
        return st

    @staticmethod
    def _layout(t: torch.Tensor):
        """(inner, row stride) of a dense tensor or a row-strided 2-D view; None if neither."""
        if t.is_contiguous():
            return t.numel(), t.numel()
        if t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]:
            return t.shape[1], t.stride(0)
        return None

    @property
    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm seen by the last step (0-dim device tensor; valid when clipping is on)."""
        return self._dev_scalars[1].sqrt() if self._dev_scalars is not None else torch.zeros(())

    def state_dict(self):
        sd = super().state_dict()
        step = float(self._dev_scalars[0].item()) if self._dev_scalars is not None else 0.0
        for st in sd["state"].values():
            st["step"] = torch.tensor(step)
        sd["fused_step"] = step
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        step = state_dict.pop("fused_step", None)
        super().load_state_dict(state_dict)
        # torch casts every loaded state tensor to the parameter dtype (bf16): restore master / moments in their own
        from itertools import chain
        saved_ids = chain.from_iterable(g["params"] for g in state_dict["param_groups"])
        mine = chain.from_iterable(g["params"] for g in self.param_groups)
        for sid, p in zip(saved_ids, mine):
            src = state_dict["state"].get(sid)
            if src is None:
                continue
            for k, dt in (("exp_avg", self.state_dtype), ("exp_avg_sq", self.state_dtype), ("master", torch.float32)):
                if k in src and torch.is_tensor(src[k]):
                    self.state[p][k] = src[k].detach().to(device=p.device, dtype=dt).clone().contiguous()
            if not self.master_weights:
                self.state[p].pop("master", None)
        if step is None:
            steps = [float(s["step"]) for s in self.state.values() if "step" in s]
            step = max(steps) if steps else 0.0
        for st in self.state.values():
            st.pop("step", None)
        dev = next((p.device for g in self.param_groups for p in g["params"]), torch.device("cpu"))
        self._scalars(dev)[0] = float(step)
        self._table_key = None

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None, max_grad_norm: Optional[float] = None):
        if closure is not None:
            raise NotImplementedError("FusedClipAdamW does not take a closure")
        max_norm = self.max_grad_norm if max_grad_norm is None else max_grad_norm
        ps, gs, ms, vs, ws, ns, inn, pst, gst, lrs, wds, updated = [], [], [], [], [], [], [], [], [], [], [], []
        device = None
        self._keepalive = []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.bfloat16 or p.grad.dtype != torch.bfloat16:
                    raise RuntimeError("FusedClipAdamW needs CUDA bf16 parameters and gradients (no CPU fallback)")
                lay = self._layout(p.data)
                if lay is None:
                    raise RuntimeError("FusedClipAdamW needs dense or row-strided 2-D parameters")
                g = p.grad
                glay = self._layout(g)
                if glay is None or glay[0] != lay[0]:
                    g = g.contiguous()
                    glay = (lay[0], lay[0]) if lay[0] != p.numel() else (p.numel(), p.numel())
                    self._keepalive.append(g)
                st = self._init_state(p)
                device = p.device
                updated.append(p)
                ps.append(p.data_ptr()); gs.append(g.data_ptr()); ms.append(st["exp_avg"].data_ptr())
                vs.append(st["exp_avg_sq"].data_ptr()); ns.append(p.numel())
                ws.append(st["master"].data_ptr() if self.master_weights else 0)
                inn.append(lay[0]); pst.append(lay[1]); gst.append(glay[1])
                lrs.append(float(group["lr"])); wds.append(float(group["weight_decay"]))
        if device is None:
            return None
        n = len(ps)
        key = (tuple(ps), tuple(gs), tuple(ms), tuple(vs), tuple(ws), tuple(pst), tuple(gst), tuple(lrs), tuple(wds))
        if key != self._table_key:
            VP, I64, F32 = ctypes.c_void_p * n, ctypes.c_int64 * n, ctypes.c_float * n
            self._tables = (VP(*ps), VP(*gs), VP(*ms), VP(*vs), VP(*ws), I64(*ns), I64(*inn), I64(*pst), I64(*gst),
                            F32(*lrs), F32(*wds))
            self._table_key = key
        sc = self._scalars(device)
        beta1, beta2 = self.param_groups[0]["betas"]
        lib = _lib.load()
        t = self._tables
        _lib.check(lib.csm_adamw_clip_step_v2(t[0], t[1], t[2], t[3], t[4] if self.master_weights else None, t[5], t[6],
                                              t[7], t[8], t[9], t[10], n, float(beta1), float(beta2),
                                              float(self.param_groups[0]["eps"]),
                                              float(max_norm) if max_norm and max_norm > 0 else 0.0,
                                              1 if self.state_dtype == torch.float32 else 0,
                                              sc.data_ptr(), sc.data_ptr() + 4,
                                              torch.cuda.current_stream().cuda_stream), "adamw_clip_step")
        self._keepalive = []
        # the kernels wrote through raw pointers: tell autograd / version-keyed caches the parameters changed
        torch.autograd.graph.increment_version(updated)
        return None
