#!/usr/bin/env python
"""Tiny launch targets for `ncu --set full` (one kernel family per invocation, a handful of launches).
   python tools/ncu_target.py gemm|ce|embed|attn|swiglu|skinny|rmsnorm"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "csm-train-pytorch_b200"))
import torch  # noqa: E402

from csm import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
what = sys.argv[1] if len(sys.argv) > 1 else "gemm"
if os.environ.get("CSM_PAIR_MODE"):            # A/B: 0 = never use the CTA-pair GEMM, 1 = whenever legal
    ops.set_gemm_cta_pair_mode(int(os.environ["CSM_PAIR_MODE"]))
if os.environ.get("CSM_NARROW_TAIL"):         # A/B: 1 = narrow MMAs on ragged last column tiles (experimental)
    ops.set_gemm_narrow_tail_mode(int(os.environ["CSM_NARROW_TAIL"]))
torch.manual_seed(0)
if what == "gemm":          # the MLP gate/up forward GEMM of one backbone layer (fused w1|w3): 4096 x 16384 x 2048
    x = torch.randn(4096, 2048, device=dev).to(BF)
    w = torch.randn(16384, 2048, device=dev).to(BF)
    out = torch.empty(4096, 16384, dtype=BF, device=dev)
    for _ in range(4):
        ops.gemm(x, w, out=out)
elif what == "ce":          # grouped audio_head fused CE forward at the c2 decoder size (232 frames)
    Dd, V, G, Ns = 1024, 2051, 31, 232
    head_t = (torch.randn(G, V, Dd, device=dev) * 0.05).to(BF)
    y = torch.randn(Ns, 32, Dd, device=dev).to(BF)
    codes = torch.randint(0, V, (Ns, 32), device=dev)
    for _ in range(4):
        ops.linear_ce_fwd(y[:, 1:], head_t, codes[:, 1:], groups=G, tgt_row_stride=32, tgt_group_stride=1)
elif what == "embed":       # K1 at the c2 size: 4096 audio frames
    C, V, Vt, D, N = 32, 2051, 128256, 2048, 4096
    audio = torch.randn(C * V, D, device=dev).to(BF)
    text = torch.randn(Vt, D, device=dev).to(BF)
    tok = torch.randint(0, V, (1, N, C + 1), device=dev)
    msk = torch.ones(1, N, C + 1, dtype=torch.bool, device=dev)
    msk[..., C] = False
    for _ in range(4):
        ops.embed_gather_sum(tok, msk, audio, text)
elif what == "swiglu":      # fused w1|w3 GEMM + SwiGLU forward and w2 dgrad + SwiGLU backward of one backbone layer
    M, I, K = 4096, 8192, 2048
    x = torch.randn(M, K, device=dev).to(BF)
    w13 = (torch.randn(2 * I, K, device=dev) * K ** -0.5).to(BF)
    w2 = (torch.randn(K, I, device=dev) * I ** -0.5).to(BF)
    dy = torch.randn(M, K, device=dev).to(BF)
    for _ in range(2):
        gu, act = ops.gemm_swiglu_fwd(x, w13)
        ops.gemm_swiglu_bwd(dy, w2, gu)
        ops.gemm(x, w13)
elif what == "attn":
    B, S, H, KV, hd = 2, 2048, 32, 8, 64
    q = torch.randn(B * S, H * hd, device=dev).to(BF)
    k = torch.randn(B * S, KV * hd, device=dev).to(BF)
    v = torch.randn(B * S, KV * hd, device=dev).to(BF)
    do = torch.randn(B * S, H * hd, device=dev).to(BF)
    for _ in range(2):
        o, lse = ops.attention_fwd(q, k, v, B, S, H, KV, hd)
        ops.attention_bwd(q, k, v, o, lse, do, B, S, H, KV, hd)
elif what == "skinny":      # the four tall-skinny LoRA products of one backbone layer (q/v adapters, r = 8 each)
    N, D, R = 4096, 2048, 16
    x = torch.randn(N, D, device=dev).to(BF)
    dy = torch.randn(N, 3072, device=dev).to(BF)
    A = torch.randn(R, D, device=dev).to(BF)
    Bm = torch.randn(3072, R, device=dev).to(BF)
    t = torch.randn(N, R, device=dev).to(BF)
    for _ in range(2):
        ops.gemm(x, A, alpha=2.0)
        ops.gemm(dy, Bm, trans_b=True, alpha=2.0)
        ops.gemm(dy, t, trans_a=True, trans_b=True)
        ops.gemm(t, x, trans_a=True, trans_b=True)
elif what == "rmsnorm":     # RMSNorm forward / backward on the fp32 residual stream, without and with the scale gradient
    N, D = 4096, 2048
    xf = torch.randn(N, D, device=dev)
    scale = torch.ones(D, device=dev, dtype=BF)
    dyn = torch.randn(N, D, device=dev).to(BF)
    dres = torch.randn(N, D, device=dev).to(BF)
    ds = torch.zeros(D, device=dev)
    for _ in range(2):
        _, rstd = ops.rmsnorm(xf, scale, 1e-5)
        ops.rmsnorm_bwd(dyn, xf, scale, rstd, dres, None)
        ops.rmsnorm_bwd(dyn, xf, scale, rstd, dres, ds)
torch.cuda.synchronize()
print("done", what)
