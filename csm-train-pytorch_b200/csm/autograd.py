"""autograd.Function wrappers: one node per *stage* of the training step (embedding, a whole transformer stack,
projection, fused linear+CE), each with a hand-written backward that calls the C-ABI kernels.

Coarse nodes keep the autograd graph to ~8 nodes per step (so the step is launch-bound on the GPU, not on Python)
and let the backward reuse exactly the activations the forward kernels wrote.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch.autograd import Function

from . import ops

BF16 = torch.bfloat16


# ----------------------------------------------------------------------------- A2: embedding gather-sum
class EmbedGatherSumFn(Function):
    """model.py:206-217 + utils.py:85-87 in one kernel."""

    @staticmethod
    def forward(ctx, tokens, mask, audio_w, text_w):
        ctx.save_for_backward(tokens, mask)
        ctx.shapes = (audio_w.shape, text_w.shape, tokens.shape[-1] - 1)
        return ops.embed_gather_sum(tokens, mask, audio_w, text_w)

    @staticmethod
    def backward(ctx, dh):
        tokens, mask = ctx.saved_tensors
        ashape, tshape, C = ctx.shapes
        da = torch.zeros(ashape, dtype=BF16, device=dh.device) if ctx.needs_input_grad[2] else None
        dt = torch.zeros(tshape, dtype=BF16, device=dh.device) if ctx.needs_input_grad[3] else None
        if da is not None or dt is not None:
            ops.embed_gather_sum_bwd(tokens, mask, dh.contiguous(), da, dt, ashape[0] // C, tshape[0])
        return None, None, da, dt


# ----------------------------------------------------------------------------- A7: decoder input gather
class DecoderInputFn(Function):
    """x[f] = [h[b,p], emb(0,c0) .. emb(C-2,c_{C-2})] — teacher-forced model.py:176,189-191."""

    @staticmethod
    def forward(ctx, h, audio_w, targets, frame_idx, codebooks: int, audio_vocab: int):
        ctx.save_for_backward(targets, frame_idx)
        ctx.meta = (h.shape, audio_w.shape, codebooks, audio_vocab)
        return ops.decoder_input(h.contiguous(), audio_w, targets, frame_idx, codebooks, audio_vocab)

    @staticmethod
    def backward(ctx, dx):
        targets, frame_idx = ctx.saved_tensors
        hshape, ashape, C, V = ctx.meta
        dh = torch.zeros(hshape, dtype=BF16, device=dx.device)
        da = torch.zeros(ashape, dtype=BF16, device=dx.device) if ctx.needs_input_grad[1] else None
        ops.decoder_input_bwd(dx, targets, frame_idx, dh, da, C, V)
        return dh, da, None, None, None, None


# ----------------------------------------------------------------------------- plain linear (projection)
class LinearFn(Function):
    @staticmethod
    def forward(ctx, x2d, w):
        ctx.save_for_backward(x2d, w)
        return ops.gemm(x2d, w)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = ops.gemm(dy, w, trans_b=True) if ctx.needs_input_grad[0] else None
        dw = ops.gemm(dy, x, trans_a=True, trans_b=True) if ctx.needs_input_grad[1] else None
        return dx, dw


# ----------------------------------------------------------------------------- transformer stack
class _Lin:
    """A (possibly LoRA-adapted) bias-free projection inside the stack: y = x W^T + t B^T, t = s * x A^T."""

    __slots__ = ("w", "A", "B", "s", "iw", "iA", "iB")

    def __init__(self, mod, index_of):
        self.w = mod.weight
        self.A = getattr(mod, "lora_A", None)
        self.B = getattr(mod, "lora_B", None)
        self.s = float(getattr(mod, "lora_scaling", 1.0))
        self.iw = index_of[id(self.w)]
        self.iA = index_of[id(self.A)] if self.A is not None else -1
        self.iB = index_of[id(self.B)] if self.B is not None else -1

    def fwd(self, x, residual=None):
        if self.A is None:
            return ops.gemm(x, self.w, residual=residual), None
        t = ops.gemm(x, self.A, alpha=self.s)                      # [N, r]
        return ops.gemm(x, self.w, residual=residual, a2=t, b2=self.B), t

    def bwd(self, dy, x, t, grads, need, dx_out=None, accumulate=False):
        """Accumulates parameter grads into `grads` and returns dx (optionally accumulated into dx_out)."""
        if need[self.iw]:
            _acc(grads, self.iw, ops.gemm(dy, x, trans_a=True, trans_b=True))
        if self.A is None:
            return ops.gemm(dy, self.w, trans_b=True, out=dx_out, accumulate=accumulate)
        dts = ops.gemm(dy, self.B, trans_b=True, alpha=self.s)     # s * dy B  [N, r]
        if need[self.iB]:
            _acc(grads, self.iB, ops.gemm(dy, t, trans_a=True, trans_b=True))      # dy^T t   [out, r]
        if need[self.iA]:
            _acc(grads, self.iA, ops.gemm(dts, x, trans_a=True, trans_b=True))     # dts^T x  [r, in]
        return ops.gemm(dy, self.w, trans_b=True, a2=dts, b2=self.A, out=dx_out, accumulate=accumulate)


def _acc(grads, i, g):
    grads[i] = g if grads[i] is None else ops.add_bf16(grads[i], g)


class StackFn(Function):
    """torchtune TransformerDecoder body (layers + final norm) — forward saves what backward needs, nothing more."""

    @staticmethod
    def forward(ctx, x, stack, *params):
        B, S, D = x.shape
        N = B * S
        H, KV, hd, eps = stack.num_heads, stack.num_kv_heads, stack.head_dim, stack.norm_eps
        if S > stack.max_seq_len:
            raise ValueError(f"seq_len ({S}) of input tensor should be smaller than max_seq_len ({stack.max_seq_len})")
        cache = stack.rope_cache(x.device)
        index_of = {id(p): i for i, p in enumerate(params)}
        cur = x.reshape(N, D).contiguous()
        saved = []
        for layer in stack.layers:
            a = layer.attn
            lq, lk, lv, lo = (_Lin(m, index_of) for m in (a.q_proj, a.k_proj, a.v_proj, a.output_proj))
            l1, l3, l2 = (_Lin(m, index_of) for m in (layer.mlp.w1, layer.mlp.w3, layer.mlp.w2))
            xn, rstd1 = ops.rmsnorm(cur, layer.sa_norm.scale, eps)
            q, tq = lq.fwd(xn)
            k, tk = lk.fwd(xn)
            v, tv = lv.fwd(xn)
            ops.rope_(q, cache, S, H, hd)
            ops.rope_(k, cache, S, KV, hd)
            o, lse = ops.attention_fwd(q, k, v, B, S, H, KV, hd)
            h, to = lo.fwd(o, residual=cur)
            hn, rstd2 = ops.rmsnorm(h, layer.mlp_norm.scale, eps)
            g, t1 = l1.fwd(hn)
            u, t3 = l3.fwd(hn)
            act = ops.swiglu(g, u)
            out, t2 = l2.fwd(act, residual=h)
            saved.append((cur, rstd1, xn, q, k, v, o, lse, h, rstd2, hn, g, u, act, (tq, tk, tv, to, t1, t3, t2),
                          (lq, lk, lv, lo, l1, l3, l2), layer))
            cur = out
        y, rstd_f = ops.rmsnorm(cur, stack.norm.scale, eps)
        ctx.saved = saved
        ctx.final = (cur, rstd_f)
        ctx.stack = stack
        ctx.index_of = index_of
        ctx.nparams = len(params)
        ctx.geom = (B, S, D)
        return y.view(B, S, D)

    @staticmethod
    def backward(ctx, dy):
        stack = ctx.stack
        B, S, D = ctx.geom
        N = B * S
        H, KV, hd = stack.num_heads, stack.num_kv_heads, stack.head_dim
        cache = stack.rope_cache(dy.device)
        need = list(ctx.needs_input_grad[2:])
        grads: List[Optional[torch.Tensor]] = [None] * ctx.nparams
        index_of = ctx.index_of
        dev = dy.device

        def norm_bwd(dyn, x, norm, rstd, dres):
            i = index_of[id(norm.scale)]
            ds = torch.zeros(x.shape[-1], dtype=torch.float32, device=dev) if need[i] else None
            dx = ops.rmsnorm_bwd(dyn, x, norm.scale, rstd, dres, ds)
            if ds is not None:
                _acc(grads, i, ds.to(BF16))
            return dx

        xf, rstd_f = ctx.final
        dcur = norm_bwd(dy.reshape(N, D).contiguous(), xf, stack.norm, rstd_f, None)
        for (x, rstd1, xn, q, k, v, o, lse, h, rstd2, hn, g, u, act, ts, lins, layer) in reversed(ctx.saved):
            tq, tk, tv, to, t1, t3, t2 = ts
            lq, lk, lv, lo, l1, l3, l2 = lins
            # ---- MLP: out = h + w2(silu(w1 hn) * w3 hn)
            dact = l2.bwd(dcur, act, t2, grads, need)
            dg, du = ops.swiglu_bwd(dact, g, u)
            dhn = l1.bwd(dg, hn, t1, grads, need)
            l3.bwd(du, hn, t3, grads, need, dx_out=dhn, accumulate=True)
            dh = norm_bwd(dhn, h, layer.mlp_norm, rstd2, dcur)
            # ---- attention: h = x + wo(attn(rope(wq xn), rope(wk xn), wv xn))
            do = lo.bwd(dh, o, to, grads, need)
            dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, do, B, S, H, KV, hd)
            ops.rope_(dq, cache, S, H, hd, inverse=True)
            ops.rope_(dk, cache, S, KV, hd, inverse=True)
            dxn = lq.bwd(dq, xn, tq, grads, need)
            lk.bwd(dk, xn, tk, grads, need, dx_out=dxn, accumulate=True)
            lv.bwd(dv, xn, tv, grads, need, dx_out=dxn, accumulate=True)
            dcur = norm_bwd(dxn, x, layer.sa_norm, rstd1, dh)
        ctx.saved = None
        dx = dcur.view(B, S, D) if ctx.needs_input_grad[0] else None
        return (dx, None) + tuple(grads)


# ----------------------------------------------------------------------------- fused linear + cross-entropy
class LinearCEFn(Function):
    """codebook0_head + F.cross_entropy(mean) (utils.py:98-107); rows with target < 0 are ignored.
    Returns (mean loss fp32 0-dim, per-row losses [M] fp32 (non-differentiable))."""

    @staticmethod
    def forward(ctx, h2d, w, targets, count: int):
        loss_rows, lse = ops.linear_ce_fwd(h2d, w, targets)
        ctx.save_for_backward(h2d, w, targets, lse)
        ctx.count = count
        loss = loss_rows.sum() / count
        rows = loss_rows[0]
        ctx.mark_non_differentiable(rows)
        return loss, rows

    @staticmethod
    def backward(ctx, g, _):
        h, w, targets, lse = ctx.saved_tensors
        dh = torch.empty_like(h)
        dw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        ops.linear_ce_bwd(h, w, targets, lse, 1.0 / ctx.count, grad_scale_dev=g.contiguous().float(), dh=dh, dw=dw)
        return dh, dw, None, None


class GroupedLinearCEFn(Function):
    """31 per-codebook heads + CE: position i of the decoder output goes through audio_head[i-1] and predicts
    code i (model.py:187).  `head_t` is the [31, V, Dd] TMA-friendly shadow of audio_head [31, Dd, V]; the gradient is
    returned in audio_head's own layout.  Returns (mean loss, per-row losses [C-1, Ns])."""

    @staticmethod
    def forward(ctx, y, head, head_t, codes):
        Ns, C, Dd = y.shape
        G = C - 1
        hv = y[:, 1:]
        tv = codes[:, 1:]
        loss_rows, lse = ops.linear_ce_fwd(hv, head_t, tv, groups=G, tgt_row_stride=codes.stride(0),
                                           tgt_group_stride=codes.stride(1))
        ctx.save_for_backward(y, head_t, codes, lse)
        loss = loss_rows.mean()
        ctx.mark_non_differentiable(loss_rows)
        return loss, loss_rows

    @staticmethod
    def backward(ctx, g, _):
        y, head_t, codes, lse = ctx.saved_tensors
        Ns, C, Dd = y.shape
        G = C - 1
        dy = torch.zeros_like(y)                     # position 0 (the backbone state) feeds no head
        dwt = torch.empty_like(head_t) if ctx.needs_input_grad[1] else None
        ops.linear_ce_bwd(y[:, 1:], head_t, codes[:, 1:], lse, 1.0 / (Ns * G), grad_scale_dev=g.contiguous().float(),
                          dh=dy[:, 1:], dw=dwt, groups=G, tgt_row_stride=codes.stride(0),
                          tgt_group_stride=codes.stride(1))
        dhead = dwt.transpose(1, 2).contiguous() if dwt is not None else None
        return dy, dhead, None, None
