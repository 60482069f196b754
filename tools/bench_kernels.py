#!/usr/bin/env python
"""Per-kernel microbenchmarks on one B200 (CUDA events, L2 evicted by reads between iterations, >=3 warm-ups).
   python tools/bench_kernels.py gemm|attn|embed|ce|all
cuBLAS / SDPA numbers are printed beside ours as same-box comparators only (never on the product path)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "csm-train-pytorch_b200"))
import torch  # noqa: E402

from csm import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
# A/B switches (defaults untouched when unset): CSM_PAIR_MODE = 0 | 1, CSM_NARROW_TAIL = 1 (experimental)
if os.environ.get("CSM_PAIR_MODE"):
    ops.set_gemm_cta_pair_mode(int(os.environ["CSM_PAIR_MODE"]))
if os.environ.get("CSM_ATTN_FWD_VARIANT"):
    ops.set_attn_fwd_variant(int(os.environ["CSM_ATTN_FWD_VARIANT"]))
if os.environ.get("CSM_DYN_TILES"):
    ops.set_gemm_dynamic_tiles(int(os.environ["CSM_DYN_TILES"]))
if os.environ.get("CSM_NARROW_TAIL"):
    ops.set_gemm_narrow_tail_mode(int(os.environ["CSM_NARROW_TAIL"]))
_flush = torch.zeros(512 << 20, dtype=torch.uint8, device=dev).view(torch.int32)


def timeit(fn, iters=10, warm=3):
    """Median CUDA-event time (ms) with a cold L2.  Two READ passes over 512 MB before every timed call: clean eviction
    (a flush by writing leaves dirty lines whose write-back competes with an HBM-bound kernel) and ~200 us during which
    the host enqueues every kernel of the op, so multi-launch ops are not timed with host launch latency inside."""
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        _flush.max()
        _flush.max()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def bench_gemm():
    shapes = [  # (M, N, K, ta, tb, tail r, label)
        (4096, 2048, 2048, 0, 0, 0, "q/o proj fwd"), (4096, 512, 2048, 0, 0, 0, "k/v proj fwd"),
        (4096, 8192, 2048, 0, 0, 0, "w1/w3 fwd"), (4096, 2048, 8192, 0, 0, 0, "w2 fwd"),
        (4096, 2048, 2048, 0, 1, 0, "q/o dgrad"), (4096, 2048, 8192, 0, 1, 0, "w1/w3 dgrad"),
        (4096, 8192, 2048, 0, 1, 0, "w2 dgrad"), (4096, 2048, 512, 0, 1, 0, "k/v dgrad"),
        (2048, 2048, 4096, 1, 1, 0, "q/o wgrad"), (8192, 2048, 4096, 1, 1, 0, "w1/w3 wgrad"),
        (4096, 2048, 2048, 0, 0, 8, "q proj fwd + LoRA tail"), (4096, 8, 2048, 0, 0, 0, "LoRA t = x A^T"),
        (2048, 8, 4096, 1, 1, 0, "LoRA dB"), (8, 2048, 4096, 1, 1, 0, "LoRA dA"),
        (7424, 1024, 2048, 0, 0, 0, "decoder projection"), (7424, 8192, 1024, 0, 0, 0, "decoder w1"),
        (8192, 8192, 8192, 0, 0, 0, "8192^3 (MEASURED_PEAKS shape)"),
    ]
    rows = []
    for M, N, K, ta, tb, r, label in shapes:
        a = torch.randn((K, M) if ta else (M, K), device=dev).to(BF)
        b = torch.randn((K, N) if tb else (N, K), device=dev).to(BF)
        a2 = torch.randn(M, r, device=dev).to(BF) if r else None
        b2 = torch.randn(N, r, device=dev).to(BF) if r else None
        out = torch.empty(M, N, dtype=BF, device=dev)
        ms = timeit(lambda: ops.gemm(a, b, trans_a=bool(ta), trans_b=bool(tb), out=out, a2=a2, b2=b2))
        A = a.t() if ta else a
        Bm = b if tb else b.t()
        ms_ref = timeit(lambda: torch.matmul(A, Bm, out=out))
        fl = 2.0 * M * N * (K + r)
        rows.append({"label": label, "M": M, "N": N, "K": K, "transA": ta, "transB": tb, "ms": ms,
                     "tflops": fl / ms / 1e9, "cublas_ms": ms_ref, "cublas_tflops": fl / ms_ref / 1e9})
        print(f"{label:32s} M={M:5d} N={N:5d} K={K:5d} ta={ta} tb={tb}: {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TF/s   "
              f"(cuBLAS {ms_ref*1e3:8.1f} us {fl/ms_ref/1e9:7.1f} TF/s)", flush=True)
    return rows


def bench_swiglu():
    rows = []
    M, I, K = 4096, 8192, 2048
    x = torch.randn(M, K, device=dev).to(BF)
    w13 = (torch.randn(2 * I, K, device=dev) * K ** -0.5).to(BF)
    w2 = (torch.randn(K, I, device=dev) * I ** -0.5).to(BF)
    dy = torch.randn(M, K, device=dev).to(BF)
    gu = ops.gemm(x, w13)
    f_fused = timeit(lambda: ops.gemm_swiglu_fwd(x, w13))
    f_gemm = timeit(lambda: ops.gemm(x, w13))
    f_sw = timeit(lambda: ops.swiglu(gu[:, :I], gu[:, I:]))
    b_fused = timeit(lambda: ops.gemm_swiglu_bwd(dy, w2, gu))
    b_gemm = timeit(lambda: ops.gemm(dy, w2, trans_b=True))
    dact = ops.gemm(dy, w2, trans_b=True)
    b_sw = timeit(lambda: ops.swiglu_bwd(dact, gu[:, :I], gu[:, I:]))
    print(f"swiglu fwd: fused {f_fused*1e3:.1f} us vs gemm {f_gemm*1e3:.1f} + swiglu {f_sw*1e3:.1f} us", flush=True)
    print(f"swiglu bwd: fused {b_fused*1e3:.1f} us vs gemm {b_gemm*1e3:.1f} + swiglu_bwd {b_sw*1e3:.1f} us", flush=True)
    rows.append({"fwd_fused_ms": f_fused, "fwd_gemm_ms": f_gemm, "fwd_swiglu_ms": f_sw, "bwd_fused_ms": b_fused,
                 "bwd_gemm_ms": b_gemm, "bwd_swiglu_ms": b_sw})
    return rows


def bench_attn():
    import torch.nn.functional as F
    rows = []
    for B, S, H, KV, hd in [(2, 2048, 32, 8, 64), (2, 4096, 32, 8, 64), (232, 32, 8, 2, 128)]:
        q = torch.randn(B * S, H * hd, device=dev).to(BF)
        k = torch.randn(B * S, KV * hd, device=dev).to(BF)
        v = torch.randn(B * S, KV * hd, device=dev).to(BF)
        do = torch.randn(B * S, H * hd, device=dev).to(BF)
        o, lse = ops.attention_fwd(q, k, v, B, S, H, KV, hd)
        f = timeit(lambda: ops.attention_fwd(q, k, v, B, S, H, KV, hd))
        bw = timeit(lambda: ops.attention_bwd(q, k, v, o, lse, do, B, S, H, KV, hd))
        fl = 4.0 * B * H * S * S * hd / 2
        q4 = q.view(B, S, H, hd).transpose(1, 2)
        k4 = k.view(B, S, KV, hd).transpose(1, 2)
        v4 = v.view(B, S, KV, hd).transpose(1, 2)
        ref = timeit(lambda: F.scaled_dot_product_attention(q4, k4, v4, is_causal=True, enable_gqa=True))
        # same-box comparators, forward AND backward (never on the product path): torch SDPA pinned to the cuDNN and to
        # the flash back-end, and the flash-attn 2 package
        comp = {}
        from torch.nn.attention import SDPBackend, sdpa_kernel
        do4 = do.view(B, S, H, hd).transpose(1, 2)
        for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("torch_flash", SDPBackend.FLASH_ATTENTION)):
            try:
                with sdpa_kernel(be):
                    qg, kg, vg = (t.detach().requires_grad_(True) for t in (q4, k4, v4))
                    tf = timeit(lambda: F.scaled_dot_product_attention(qg, kg, vg, is_causal=True, enable_gqa=True))
                    out = F.scaled_dot_product_attention(qg, kg, vg, is_causal=True, enable_gqa=True)
                    tb = timeit(lambda: torch.autograd.grad(out, (qg, kg, vg), do4, retain_graph=True))
                comp[name] = {"fwd_ms": tf, "bwd_ms": tb}
            except Exception as e:  # noqa: BLE001
                comp[name] = {"error": str(e)[:100]}
        try:
            from flash_attn import flash_attn_func
            qf, kf, vf = (t.view(B, S, n, hd).detach().requires_grad_(True) for t, n in ((q, H), (k, KV), (v, KV)))
            tf = timeit(lambda: flash_attn_func(qf, kf, vf, causal=True))
            out = flash_attn_func(qf, kf, vf, causal=True)
            dof = do.view(B, S, H, hd)
            tb = timeit(lambda: torch.autograd.grad(out, (qf, kf, vf), dof, retain_graph=True))
            comp["flash_attn_2"] = {"fwd_ms": tf, "bwd_ms": tb}
        except Exception as e:  # noqa: BLE001
            comp["flash_attn_2"] = {"error": str(e)[:100]}
        rows.append({"B": B, "S": S, "H": H, "KV": KV, "hd": hd, "fwd_ms": f, "bwd_ms": bw,
                     "fwd_tflops": fl / f / 1e9, "bwd_tflops": 2.5 * fl / bw / 1e9, "sdpa_fwd_ms": ref,
                     "comparators": comp})
        cs = "  ".join(f"{n}: fwd {c['fwd_ms']*1e3:.0f} bwd {c['bwd_ms']*1e3:.0f} us" if "fwd_ms" in c else f"{n}: n/a"
                       for n, c in comp.items())
        print(f"attn B={B} S={S} H={H} KV={KV} hd={hd}: fwd {f*1e3:.0f} us ({fl/f/1e9:.0f} TF/s causal-alg) "
              f"bwd {bw*1e3:.0f} us ({2.5*fl/bw/1e9:.0f} TF/s)  [SDPA fwd {ref*1e3:.0f} us]  [{cs}]", flush=True)
    return rows


def bench_embed():
    C, V, Vt, D = 32, 2051, 128256, 2048
    audio = torch.randn(C * V, D, device=dev).to(BF)
    text = torch.randn(Vt, D, device=dev).to(BF)
    rows = []
    for N in (4096, 65536):
        tok = torch.randint(0, V, (1, N, C + 1), device=dev)
        msk = torch.ones(1, N, C + 1, dtype=torch.bool, device=dev)
        msk[..., C] = False                     # audio frames: 32 rows each
        ms = timeit(lambda: ops.embed_gather_sum(tok, msk, audio, text))
        byts = N * (32 * D * 2 + D * 2 + 33 * 8 + 33)
        rows.append({"frames": N, "ms": ms, "GBs": byts / ms / 1e6})
        print(f"embed_gather_sum audio frames N={N}: {ms*1e3:.1f} us  {byts/ms/1e6:.0f} GB/s (algorithmic)", flush=True)
    return rows


def bench_ce():
    rows = []
    Dd, V, G = 1024, 2051, 31
    head_t = (torch.randn(G, V, Dd, device=dev) * 0.05).to(BF)
    for Ns in (64, 128, 232, 512, 1024, 2048, 4096, 8192):
        y = torch.randn(Ns, 32, Dd, device=dev).to(BF)
        codes = torch.randint(0, V, (Ns, 32), device=dev)
        fn = lambda: ops.linear_ce_fwd(y[:, 1:], head_t, codes[:, 1:], groups=G, tgt_row_stride=32, tgt_group_stride=1)
        ms = timeit(fn)
        byts = G * Dd * V * 2 + Ns * G * (Dd * 2 + 8 + 4)
        fl = 2.0 * Ns * G * Dd * V
        rows.append({"N_sel": Ns, "ms": ms, "GBs": byts / ms / 1e6, "tflops": fl / ms / 1e9})
        print(f"grouped audio_head CE fwd N_sel={Ns}: {ms*1e3:.1f} us  {byts/ms/1e6:.0f} GB/s  {fl/ms/1e9:.0f} TF/s",
              flush=True)
    M, K = 4096, 2048
    h = torch.randn(M, K, device=dev).to(BF)
    w = (torch.randn(V, K, device=dev) * 0.03).to(BF)
    t = torch.randint(0, V, (M,), device=dev)
    ms = timeit(lambda: ops.linear_ce_fwd(h, w, t))
    print(f"codebook0_head CE fwd M={M}: {ms*1e3:.1f} us {2.0*M*K*V/ms/1e9:.0f} TF/s", flush=True)
    rows.append({"c0_M": M, "ms": ms})
    return rows


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    out = {}
    if what in ("gemm", "all"):
        out["gemm"] = bench_gemm()
    if what in ("swiglu", "all"):
        out["swiglu"] = bench_swiglu()
    if what in ("attn", "all"):
        out["attn"] = bench_attn()
    if what in ("embed", "all"):
        out["embed"] = bench_embed()
    if what in ("ce", "all"):
        out["ce"] = bench_ce()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"kernels_{what}.json"), "w") as f:
        json.dump(out, f, indent=1)
