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
# The w2-dgrad GEMM applies the SwiGLU backward in its epilogue (gate/up fetched one chunk ahead as whole 64-byte row
# segments and transposed to the row-per-lane TMEM layout through smem): dact never exists and the separate swiglu_bwd
# pass (1.3 ms per c2 step) disappears.  Cold-cache microbenchmark: 120 us against 87 + 67 us at 4096 x 8192 x 2048.
# In-step A/B on B200 (round 2, same box, alternating runs): 24.33 / 24.37 ms fused vs 24.48 / 24.60 ms unfused — in
# round 1 the unfused path had won by 0.6 ms because swiglu_bwd read dact out of L2; with the fp32 residual stream and
# the other round-2 changes that no longer holds.  CSM_FUSE_SWIGLU_BWD=0 selects the unfused kernels.
import os as _os
FUSE_SWIGLU_BWD = _os.environ.get("CSM_FUSE_SWIGLU_BWD", "1") == "1"
# The residual stream of both transformer stacks (x -> h = x + attn(..) -> out = h + mlp(..)) is kept in fp32 between
# layers: the o-proj / down-proj GEMM epilogues add the fp32 residual and store fp32, RMSNorm reads fp32.  Measured
# (tools/parity_probe*.py, profiles/r2_parity_*.txt): with a bf16 stream — what stock bf16 PyTorch does — the q/k
# projection gradients of the late layers agree with the fp32 oracle only to cosine 0.9986-0.9989 (stock torch bf16:
# 0.9986), because dS = P o (dP - delta) amplifies the rounding noise accumulated in the stream; with the fp32 stream
# every trainable tensor clears the 0.999 gate.  Costs ~2 % of a step in extra bytes.  The gradient stream stays bf16.
FP32_RESIDUAL = True


# ----------------------------------------------------------------------------- A2: embedding gather-sum
class EmbedGatherSumFn(Function):
    """model.py:206-217 + utils.py:85-87 in one kernel.

    `text_exchange` (data-parallel full fine-tune, csm/training/dp.py): the text-embedding gradient touches at most one
    row per frame, so instead of all-reducing the dense [128256, 2048] table the ranks all-gather (tokens, mask, dh) and
    every rank scatters ALL ranks' rows into its own dense gradient — the same scatter kernel, ~17 MB per rank on the
    wire instead of 525 MB."""

    @staticmethod
    def forward(ctx, tokens, mask, audio_w, text_w, text_exchange=None):
        ctx.save_for_backward(tokens, mask)
        ctx.shapes = (audio_w.shape, text_w.shape, tokens.shape[-1] - 1)
        ctx.text_exchange = text_exchange
        ctx.packed = ops.is_packed_tokens(tokens, mask)     # compact device format (csm/data/frames.py::pack_tokens)
        if ctx.packed:
            return ops.embed_gather_sum_packed(tokens, mask, audio_w, text_w)
        return ops.embed_gather_sum(tokens, mask, audio_w, text_w)

    @staticmethod
    def backward(ctx, dh):
        tokens, mask = ctx.saved_tensors
        ashape, tshape, C = ctx.shapes
        dh = dh.contiguous()
        if ctx.packed:
            da = torch.zeros(ashape, dtype=BF16, device=dh.device) if ctx.needs_input_grad[2] else None
            dt = None
            if ctx.needs_input_grad[3]:
                if ctx.text_exchange is not None:
                    # the rank exchange works on the unpacked layout; only the text column matters to it
                    W = tokens.shape[-1]
                    bits = (mask.unsqueeze(-1) >> torch.arange(W, device=mask.device)) & 1
                    dt = ctx.text_exchange(tokens.to(torch.int64), bits.to(torch.uint8), dh, tshape)
                else:
                    dt = torch.zeros(tshape, dtype=BF16, device=dh.device)
            local_dt = dt if ctx.text_exchange is None else None
            if da is not None or local_dt is not None:
                ops.embed_gather_sum_packed_bwd(tokens, mask, dh, da, local_dt, ashape[0] // C, tshape[0])
            return None, None, da, dt, None
        da = torch.zeros(ashape, dtype=BF16, device=dh.device) if ctx.needs_input_grad[2] else None
        dt = None
        if ctx.needs_input_grad[3]:
            if ctx.text_exchange is not None:
                dt = ctx.text_exchange(tokens, mask, dh, tshape)        # dense, already averaged over the ranks
            else:
                dt = torch.zeros(tshape, dtype=BF16, device=dh.device)
        local_dt = dt if ctx.text_exchange is None else None
        if da is not None or local_dt is not None:
            ops.embed_gather_sum_bwd(tokens, mask, dh, da, local_dt, ashape[0] // C, tshape[0])
        return None, None, da, dt, None


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

    __slots__ = ("w", "A", "B", "s", "iw", "iA", "iB", "rank", "K", "rows")

    def __init__(self, mod, index_of, adapter_rows=None):
        self.w = mod.weight
        self.A = getattr(mod, "lora_A", None)
        self.B = getattr(mod, "lora_B", None)
        self.s = float(getattr(mod, "lora_scaling", 1.0))
        self.rank = int(getattr(mod, "lora_r", 0))
        self.K = int(getattr(mod, "lora_adapters", 1))
        self.rows = adapter_rows if self.K > 1 else None     # int32 [N]: adapter of every row (multi-adapter batching)
        if self.A is not None and self.K > 1 and adapter_rows is None:
            raise RuntimeError("this model holds several LoRA adapters per projection: pass speaker_ids")
        self.iw = index_of[id(self.w)]
        self.iA = index_of[id(self.A)] if self.A is not None else -1
        self.iB = index_of[id(self.B)] if self.B is not None else -1

    def fwd(self, x, residual=None, out_dtype=BF16):
        if self.A is None:
            return ops.gemm(x, self.w, residual=residual, out_dtype=out_dtype), None
        t = ops.gemm(x, self.A, alpha=self.s)                      # [N, K*r]
        if self.rows is not None:
            ops.lora_mask_rows_(t, self.rows, self.rank, self.K)
        return ops.gemm(x, self.w, residual=residual, a2=t, b2=self.B, out_dtype=out_dtype), t

    def bwd(self, dy, x, t, grads, need, dx_out=None, accumulate=False, swiglu_gu=None, sink=None, written=None):
        """Accumulates parameter grads into `grads` and returns dx (optionally accumulated into dx_out).
        With `swiglu_gu` (the saved gate|up buffer of a fused MLP) the dgrad GEMM's epilogue applies the SwiGLU
        backward and the return value is dgate|dup instead of dx.  With a data-parallel `sink` the wgrad GEMM writes
        straight into the all-reduce bucket (index noted in `written`)."""
        if need[self.iw]:
            dst = sink.grad_view(self.w) if sink is not None else None
            if dst is not None:
                ops.gemm(dy, x, trans_a=True, trans_b=True, out=dst, accumulate=sink.has_grad(self.w))
                written.add(self.iw)
            else:
                _acc(grads, self.iw, ops.gemm(dy, x, trans_a=True, trans_b=True))
        if self.A is None:
            if swiglu_gu is not None:
                return ops.gemm_swiglu_bwd(dy, self.w, swiglu_gu)
            return ops.gemm(dy, self.w, trans_b=True, out=dx_out, accumulate=accumulate)
        dts = ops.gemm(dy, self.B, trans_b=True, alpha=self.s)     # s * dy B  [N, K*r]
        if self.rows is not None:
            ops.lora_mask_rows_(dts, self.rows, self.rank, self.K)
        if need[self.iB]:
            _acc(grads, self.iB, ops.gemm(dy, t, trans_a=True, trans_b=True))      # dy^T t   [out, r]
        if need[self.iA]:
            _acc(grads, self.iA, ops.gemm(dts, x, trans_a=True, trans_b=True))     # dts^T x  [r, in]
        if swiglu_gu is not None:
            return ops.gemm_swiglu_bwd(dy, self.w, swiglu_gu, a2=dts, b2=self.A)
        return ops.gemm(dy, self.w, trans_b=True, a2=dts, b2=self.A, out=dx_out, accumulate=accumulate)


def _acc(grads, i, g):
    if grads[i] is None:
        grads[i] = g
    else:
        grads[i] = ops.add_bf16(grads[i].contiguous(), g.contiguous())


def _packed_weight(mods, cache):
    """One [sum(out_i), in] tensor holding the weights of several projections that share their input; the modules'
    ``weight.data`` become row-slice views of it, so optimiser updates and ``load_state_dict`` write straight into the
    fused operand.  Re-packed if the views were broken (e.g. by ``model.to(device)``)."""
    key = tuple(id(m) for m in mods)
    fused = cache.get(key)
    ok = fused is not None and fused.device == mods[0].weight.device and fused.dtype == mods[0].weight.dtype
    if ok:
        off = 0
        for m in mods:
            w = m.weight
            if w.data_ptr() != fused.data_ptr() + off * fused.stride(0) * fused.element_size() or \
                    w.stride(0) != fused.stride(0):
                ok = False
                break
            off += w.shape[0]
    if not ok:
        with torch.no_grad():
            fused = torch.cat([m.weight.detach() for m in mods], dim=0).contiguous()
            off = 0
            for m in mods:
                n = m.weight.shape[0]
                m.weight.data = fused[off:off + n]
                off += n
        cache[key] = fused
    return fused


def _packed_lora(lora, n_out, offs, W, cache, key):
    """Persistent operands of a fused LoRA group: A_cat [R, in] (the adapters' A factors stacked) and the block-diagonal
    B_bd [n_out, R]; the modules' ``lora_A`` / ``lora_B`` parameters are re-pointed ONCE to row-slice / block views of
    them (``lora_B`` becomes a row-strided view: the optimiser and the gradient exchange take strided tensors), so a
    training step issues no torch.cat / copy kernels for the low-rank operands (round 1: ~100 launches, 0.5 ms per
    c2 step).  Re-packed if the views were broken (e.g. by ``model.to(device)``)."""
    R = sum(m.lora_A.shape[0] for _, m in lora)
    packed = cache.get(key)
    ok = packed is not None and packed[0].device == W.device and packed[0].dtype == W.dtype and \
        packed[0].shape == (R, W.shape[1]) and packed[1].shape == (n_out, R)
    if ok:
        A_cat, B_bd = packed
        ro = 0
        for j, m in lora:
            r = m.lora_A.shape[0]
            if m.lora_A.data_ptr() != A_cat.data_ptr() + ro * A_cat.stride(0) * A_cat.element_size() or \
                    m.lora_B.data_ptr() != B_bd.data_ptr() + (offs[j] * R + ro) * B_bd.element_size() or \
                    m.lora_B.stride(0) != R:
                ok = False
                break
            ro += r
    if not ok:
        with torch.no_grad():
            A_cat = torch.cat([m.lora_A.detach() for _, m in lora], dim=0).contiguous()
            B_bd = torch.zeros(n_out, R, dtype=W.dtype, device=W.device)
            ro = 0
            for j, m in lora:
                r, rows = m.lora_A.shape[0], m.weight.shape[0]
                blk = B_bd[offs[j]:offs[j] + rows, ro:ro + r]
                blk.copy_(m.lora_B.detach())
                m.lora_A.data = A_cat[ro:ro + r]
                m.lora_B.data = blk
                ro += r
        cache[key] = (A_cat, B_bd)
    return A_cat, B_bd


class _Group:
    """Several bias-free projections of the SAME input as one GEMM: y = x [W_0; W_1; ..]^T  (q/k/v, gate/up).
    LoRA adapters on any member ride in the same GEMM through the extra K block: t = x [A_0; A_1; ..]^T and a
    block-diagonal [(s_0 B_0) 0; 0 (s_1 B_1)] tail operand."""

    def __init__(self, mods, index_of, cache, adapter_rows=None, drop=None):
        """`drop` = (p, seed_dev int64 [1], salt): LoRA input dropout of this call (training only; lora.py:87-90)."""
        self.mods = mods
        self.rows = None
        self.drop = drop if (drop is not None and drop[0] > 0.0) else None
        self.bias = [(j, m) for j, m in enumerate(mods) if getattr(m, "lora_bias", None) is not None]
        self.ib = {j: index_of[id(m.lora_bias)] for j, m in self.bias}
        self.W = _packed_weight(mods, cache)
        self.iw = [index_of[id(m.weight)] for m in mods]
        self.offs, o = [], 0
        for m in mods:
            self.offs.append(o)
            o += m.weight.shape[0]
        self.n_out = o
        self.lora = [(j, m) for j, m in enumerate(mods) if getattr(m, "lora_A", None) is not None]
        self.A_cat = self.B_bd = None
        if self.lora:
            R = sum(m.lora_A.shape[0] for _, m in self.lora)
            if R > 256:
                raise RuntimeError("fused LoRA group: total rank (projections x adapters x r) must be <= 256")
            Ks = {int(getattr(m, "lora_adapters", 1)) for _, m in self.lora}
            rs = {int(m.lora_r) for _, m in self.lora}
            if max(Ks) > 1:
                if len(Ks) != 1 or len(rs) != 1:
                    raise RuntimeError("multi-adapter LoRA: every projection of a fused group needs the same r and count")
                if adapter_rows is None:
                    raise RuntimeError("this model holds several LoRA adapters per projection: pass speaker_ids")
                self.rows, self.rank, self.K = adapter_rows, rs.pop(), Ks.pop()
            # the common scaling alpha/r rides on the skinny GEMMs' alpha (t = s x A^T, dts = s dy B): the block-diagonal
            # tail operand then holds the raw B factors and no per-step scaling kernels are needed
            scal = {float(m.lora_scaling) for _, m in self.lora}
            self.s = scal.pop() if len(scal) == 1 else None
            self.R = R
            # common scaling: the parameters can BE the operands (not with the LoRA bias column, which is appended)
            self.views = self.s is not None and R % 8 == 0 and not self.bias
            if self.views:
                self.A_cat, self.B_bd = _packed_lora(self.lora, self.n_out, self.offs, self.W, cache,
                                                     ("lora",) + tuple(id(m) for m in mods))
            else:
                self.A_cat = torch.cat([m.lora_A.detach() for _, m in self.lora], dim=0)        # [R, in]
                key = ("B_bd",) + tuple(id(m) for m in mods)
                self.B_bd = cache.get(key)
                if self.B_bd is None or self.B_bd.shape != (self.n_out, R) or self.B_bd.device != self.W.device:
                    self.B_bd = cache[key] = torch.zeros(self.n_out, R, dtype=self.W.dtype, device=self.W.device)
            self.slots, ro = [], 0
            for j, m in self.lora:
                r = m.lora_A.shape[0]
                if not self.views:
                    blk = self.B_bd[self.offs[j]:self.offs[j] + m.weight.shape[0], ro:ro + r]
                    if self.s is None:
                        torch.mul(m.lora_B.detach(), float(m.lora_scaling), out=blk)
                    else:
                        blk.copy_(m.lora_B.detach())
                self.slots.append((j, m, ro, r, index_of[id(m.lora_A)], index_of[id(m.lora_B)]))
                ro += r

    def _xd(self, x):
        """The input of the low-rank path: x, or dropout(x) in training (mask re-derived from the seed in backward)."""
        if self.drop is None:
            return x
        p, seed, salt = self.drop
        return ops.lora_dropout(x, p, seed, salt)

    def _t(self, x):
        s = self.s if self.s is not None else 1.0
        if not self.bias:
            t = ops.gemm(self._xd(x), self.A_cat, alpha=s)                                       # [N, R]
            if self.rows is not None:
                ops.lora_mask_rows_(t, self.rows, self.rank, self.K)
            return t
        # LoRA bias (lora.py:66,101-102: y += lora_bias, unscaled): one more tail column — t gets a column of ones and
        # the tail operand the bias values, so the bias add rides in the same GEMM and its gradient (column sums of dy)
        # falls out of the dB GEMM.  Eight columns are appended to keep the 16-byte row alignment.
        if self.rows is not None:
            raise RuntimeError("lora_use_bias is not supported together with several adapters per projection")
        R = self.R
        t = torch.zeros(x.shape[0], R + 8, dtype=x.dtype, device=x.device)
        ops.gemm(self._xd(x), self.A_cat, alpha=s, out=t[:, :R])
        t[:, R].fill_(1.0)
        Bx = torch.zeros(self.n_out, R + 8, dtype=x.dtype, device=x.device)
        Bx[:, :R].copy_(self.B_bd)
        for j, m in self.bias:
            Bx[self.offs[j]:self.offs[j] + m.weight.shape[0], R].copy_(m.lora_bias.detach())
        self.B_ext = Bx
        return t

    @property
    def b2(self):
        return self.B_ext if self.bias else self.B_bd

    def fwd(self, x, rope=None, residual=None, out_dtype=BF16):
        """rope = (cache, seq_len, rope_cols, head_dim[, positions int32 [N]]): rotate the leading columns in the GEMM's
        store epilogue (positions: sequence packing, they restart with every packed sample)."""
        t = self._t(x) if self.lora else None
        b2 = self.b2 if self.lora else None
        if rope is not None:
            return ops.gemm_rope(x, self.W, *rope[:4], a2=t, b2=b2, positions=rope[4] if len(rope) > 4 else None), t
        return ops.gemm(x, self.W, a2=t, b2=b2, residual=residual, out_dtype=out_dtype), t

    def fwd_swiglu(self, x):
        """gate/up group only: (gate|up, act = silu(gate) * up, t) with the SwiGLU in the GEMM epilogue."""
        if not self.lora:
            gu, act = ops.gemm_swiglu_fwd(x, self.W)
            return gu, act, None
        t = self._t(x)
        gu, act = ops.gemm_swiglu_fwd(x, self.W, a2=t, b2=self.b2)
        return gu, act, t

    def bwd(self, dy, x, t, grads, need, sink=None, written=None, need_dx=True):
        if any(need[i] for i in self.iw):
            ws = [m.weight for m in self.mods]
            dst = sink.packed_view(ws) if sink is not None and all(need[i] for i in self.iw) else None
            if dst is not None:
                # the whole [n_out, in] weight gradient of the group lands in the data-parallel bucket: no copy
                ops.gemm(dy, x, trans_a=True, trans_b=True, out=dst, accumulate=sink.has_grad(ws[0]))
                written.update(self.iw)
            else:
                dW = ops.gemm(dy, x, trans_a=True, trans_b=True)          # [n_out, in] in one GEMM
                for j, m in enumerate(self.mods):
                    if need[self.iw[j]]:
                        _acc(grads, self.iw[j], dW[self.offs[j]:self.offs[j] + m.weight.shape[0]])
        if not self.lora:
            return ops.gemm(dy, self.W, trans_b=True) if need_dx else None
        # with the common scaling s folded into t and dts:  y = x W^T + t B^T,  t = s x A^T
        #   dB = dy^T t,   dts = s dy B,   dA = dts^T x,   dx = dy W + dts A
        s = self.s if self.s is not None else 1.0
        dts = ops.gemm(dy, self.b2, trans_b=True, alpha=s)             # [N, R (+8 with the bias column)]
        if self.bias:
            dts = dts[:, :self.R]                                      # the ones column is not a function of x
        if self.rows is not None:
            ops.lora_mask_rows_(dts, self.rows, self.rank, self.K)
        if any(need[iB] for *_, iB in self.slots) or self.bias:
            dB = ops.gemm(dy, t, trans_a=True, trans_b=True)          # dy^T t       [n_out, R (+8)]
            for j, m in self.bias:
                if need[self.ib[j]]:
                    _acc(grads, self.ib[j], dB[self.offs[j]:self.offs[j] + m.weight.shape[0], self.R].contiguous())
        xd = self._xd(x)                                              # same seed -> same dropout mask as the forward
        if any(need[iA] for *_, iA, _ in self.slots):
            dA = ops.gemm(dts, xd, trans_a=True, trans_b=True)        # dts^T x      [R, in]
        for j, m, ro, r, iA, iB in self.slots:
            if need[iB]:
                blk = dB[self.offs[j]:self.offs[j] + m.weight.shape[0], ro:ro + r]
                if self.views and m.lora_B.grad is None and not torch.is_grad_enabled():
                    # row-strided like the parameter itself.  Handed to the parameter directly: returned through
                    # autograd, AccumulateGrad would clone it into a dense tensor (one copy kernel per adapter)
                    m.lora_B.grad = blk
                elif self.views:
                    _acc(grads, iB, blk)
                else:
                    _acc(grads, iB, blk.contiguous() if self.s is not None else
                         (blk * float(m.lora_scaling)).contiguous())
            if need[iA]:
                _acc(grads, iA, dA[ro:ro + r])
        if not need_dx:
            return None
        if self.drop is not None:
            # dx = dy W + dropout_mask o (dts A) / (1 - p): the low-rank term passes back through the input mask
            dx = ops.gemm(dy, self.W, trans_b=True)
            tmp = ops.gemm(dts, self.A_cat, trans_b=True)
            p_, seed, salt = self.drop
            return ops.lora_dropout(tmp, p_, seed, salt, out=dx, accumulate=True)
        return ops.gemm(dy, self.W, trans_b=True, a2=dts, b2=self.A_cat)


class StackFn(Function):
    """torchtune TransformerDecoder body (layers + final norm) — forward saves what backward needs, nothing more.
    Per layer: 4 GEMMs forward (fused qkv, o, fused gate/up, down), residual adds in the GEMM epilogues."""

    @staticmethod
    def forward(ctx, x, stack, *params):
        B, S, D = x.shape
        N = B * S
        H, KV, hd, eps = stack.num_heads, stack.num_kv_heads, stack.head_dim, stack.norm_eps
        if S > stack.max_seq_len:
            raise ValueError(f"seq_len ({S}) of input tensor should be smaller than max_seq_len ({stack.max_seq_len})")
        cache = stack.rope_cache(x.device)
        index_of = {id(p): i for i, p in enumerate(params)}
        nq, nkv = H * hd, KV * hd
        rows = getattr(stack, "_adapter_rows", None)              # multi-adapter LoRA: int32 [N] adapter of each row
        if rows is not None and rows.numel() != N:
            raise RuntimeError(f"adapter row ids: expected {N} entries, got {rows.numel()}")
        # sequence packing: (seg_start int32 [B,S], seg_end int32 [B,S], positions int32 [B*S]) — block-diagonal causal
        # attention and RoPE positions that restart with every packed sample
        packing = getattr(stack, "_packing", None)
        seg_start, seg_end, pos_rows = packing if packing is not None else (None, None, None)
        cur = x.reshape(N, D).contiguous()
        res_dtype = torch.float32 if FP32_RESIDUAL else BF16       # dtype of h / out (cur is bf16 for layer 0 only)
        saved = []
        # LoRA options of the reference's LoRALinear (lora.py:87-90,101-102): input dropout (training only) and a bias
        p_drop = float(getattr(stack, "lora_dropout", 0.0)) if stack.training else 0.0
        seed = getattr(stack, "_lora_seed", None)
        if p_drop > 0.0 and (seed is None or seed.device != x.device):
            seed = stack._lora_seed = torch.zeros(1, dtype=torch.int64, device=x.device)
        if p_drop > 0.0:
            seed.add_(1)                                           # a fresh mask every step, also under graph replay

        def lin(mod, salt):
            if getattr(mod, "lora_A", None) is not None and (p_drop > 0.0 or getattr(mod, "lora_bias", None) is not None):
                return _Group([mod], index_of, stack._packed, rows, (p_drop, seed, salt))
            return _Lin(mod, index_of, rows)

        for li, layer in enumerate(stack.layers):
            a = layer.attn
            gqkv = _Group([a.q_proj, a.k_proj, a.v_proj], index_of, stack._packed, rows, (p_drop, seed, 4 * li))
            g13 = _Group([layer.mlp.w1, layer.mlp.w3], index_of, stack._packed, rows, (p_drop, seed, 4 * li + 1))
            lo, l2 = lin(a.output_proj, 4 * li + 2), lin(layer.mlp.w2, 4 * li + 3)
            I = layer.mlp.w1.weight.shape[0]
            xn, rstd1 = ops.rmsnorm(cur, layer.sa_norm.scale, eps)
            qkv, tqkv = gqkv.fwd(xn, rope=(cache, S, nq + nkv, hd, pos_rows))   # q, k rotated in the store epilogue
            q, k, v = qkv[:, :nq], qkv[:, nq:nq + nkv], qkv[:, nq + nkv:]
            o, lse = ops.attention_fwd(q, k, v, B, S, H, KV, hd, seg_start=seg_start)
            h, to = lo.fwd(o, residual=cur, out_dtype=res_dtype)
            hn, rstd2 = ops.rmsnorm(h, layer.mlp_norm.scale, eps)
            if ops.swiglu_fusable(N, I, D):
                gu, act, t13 = g13.fwd_swiglu(hn)
            else:
                gu, t13 = g13.fwd(hn)
                act = ops.swiglu(gu[:, :I], gu[:, I:])
            out, t2 = l2.fwd(act, residual=h, out_dtype=res_dtype)
            saved.append((cur, rstd1, xn, qkv, o, lse, h, rstd2, hn, gu, act, (tqkv, to, t13, t2),
                          (gqkv, lo, g13, l2), layer))
            cur = out
        y, rstd_f = ops.rmsnorm(cur, stack.norm.scale, eps)
        ctx.saved = saved
        ctx.final = (cur, rstd_f)
        ctx.stack = stack
        ctx.index_of = index_of
        ctx.nparams = len(params)
        ctx.geom = (B, S, D)
        ctx.packing = (seg_start, seg_end)
        return y.view(B, S, D)

    @staticmethod
    def backward(ctx, dy):
        stack = ctx.stack
        B, S, D = ctx.geom
        N = B * S
        H, KV, hd = stack.num_heads, stack.num_kv_heads, stack.head_dim
        nq, nkv = H * hd, KV * hd
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

        sink = getattr(stack, "_grad_sink", None)   # data-parallel bucket owner (csm/training/dp.py), optional
        written = set()                             # parameter indices whose wgrad GEMM wrote straight into its bucket

        def hand_over(module):
            # gradients of `module` are final: give them to the DP synchroniser now so their all-reduce overlaps the
            # rest of this backward; autograd then receives None for them
            if sink is None:
                return
            for prm in module.parameters():
                i = index_of[id(prm)]
                if i in written:
                    sink.mark_written(prm)
                elif grads[i] is not None and sink.deliver(prm, grads[i]):
                    grads[i] = None

        xf, rstd_f = ctx.final
        dcur = norm_bwd(dy.reshape(N, D).contiguous(), xf, stack.norm, rstd_f, None)
        hand_over(stack.norm)
        for (x, rstd1, xn, qkv, o, lse, h, rstd2, hn, gu, act, ts, lins, layer) in reversed(ctx.saved):
            tqkv, to, t13, t2 = ts
            gqkv, lo, g13, l2 = lins
            I = gu.shape[1] // 2
            # ---- MLP: out = h + w2(silu(w1 hn) * w3 hn)
            if FUSE_SWIGLU_BWD and isinstance(l2, _Lin) and ops.swiglu_fusable(N, I, D):
                dgu = l2.bwd(dcur, act, t2, grads, need, swiglu_gu=gu, sink=sink, written=written)
            else:
                dact = l2.bwd(dcur, act, t2, grads, need, sink=sink, written=written)
                dgu = torch.empty_like(gu)
                ops.swiglu_bwd(dact, gu[:, :I], gu[:, I:], dgate=dgu[:, :I], dup=dgu[:, I:])
            dhn = g13.bwd(dgu, hn, t13, grads, need, sink=sink, written=written)
            dh = norm_bwd(dhn, h, layer.mlp_norm, rstd2, dcur)
            # ---- attention: h = x + wo(attn(rope(wq xn), rope(wk xn), wv xn))
            do = lo.bwd(dh, o, to, grads, need, sink=sink, written=written)
            q, k, v = qkv[:, :nq], qkv[:, nq:nq + nkv], qkv[:, nq + nkv:]
            dqkv = torch.empty_like(qkv)
            dq, dk, dv = dqkv[:, :nq], dqkv[:, nq:nq + nkv], dqkv[:, nq + nkv:]
            # (the inverse RoPE of dq / dk happens in the attention kernels' store epilogues when they can)
            ops.attention_bwd(q, k, v, o, lse, do, B, S, H, KV, hd, dq=dq, dk=dk, dv=dv, rope_cache=cache,
                              seg_start=ctx.packing[0], seg_end=ctx.packing[1])
            first = layer is stack.layers[0]
            if first and not ctx.needs_input_grad[0] and not need[index_of[id(layer.sa_norm.scale)]]:
                # LoRA on frozen embeddings: nobody consumes the gradient of the stack's input — the first layer's
                # q|k|v dgrad GEMM and its RMSNorm backward are skipped (the adapters' own gradients are still taken)
                gqkv.bwd(dqkv, xn, tqkv, grads, need, sink=sink, written=written, need_dx=False)
                dcur = None
            else:
                dxn = gqkv.bwd(dqkv, xn, tqkv, grads, need, sink=sink, written=written)
                dcur = norm_bwd(dxn, x, layer.sa_norm, rstd1, dh)
            hand_over(layer)
        ctx.saved = None
        dx = dcur.view(B, S, D) if (ctx.needs_input_grad[0] and dcur is not None) else None
        return (dx, None) + tuple(grads)


# ----------------------------------------------------------------------------- fused linear + cross-entropy
class LinearCEFn(Function):
    """codebook0_head + F.cross_entropy(mean) (utils.py:98-107); rows with target < 0 are ignored.
    Returns (mean loss fp32 0-dim, per-row losses [M] fp32 (non-differentiable))."""

    @staticmethod
    def forward(ctx, h2d, w, targets, count):
        """`count`: rows the mean runs over — a Python int, or a 0-dim device tensor when it depends on the batch."""
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
        g = g.contiguous().float()
        if torch.is_tensor(ctx.count):
            scale, g = 1.0, g / ctx.count
        else:
            scale = 1.0 / ctx.count
        ops.linear_ce_bwd(h, w, targets, lse, scale, grad_scale_dev=g, dh=dh, dw=dw)
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
