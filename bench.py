#!/usr/bin/env python
"""CSM-1B train-step throughput (audio frames/s) on N B200s — BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W            # this repo (libcsm_b200 kernels)
    python bench.py --impl reference --steps K --warmup W    # the reference path on the host CPU cores

A "step" = forward + backward + gradient all-reduce (N>1) + AdamW of the CSM training step on one synthetic batch.
Default workload = BASELINE.json configs[1]: CSM-1B (Llama-3.2-1B backbone + 100M decoder), LoRA r=8 on q_proj/v_proj,
bf16, 2048-frame sequences, decoder trained on 1/16 of the audio frames, batch 2 per GPU.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "csm-train-pytorch_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# NCCL writes its banner / debug lines to STDOUT when NCCL_DEBUG is VERSION or above (what some launchers export);
# the driver reads one JSON line from stdout, so NCCL's output goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import torch  # noqa: E402

METRIC = "CSM-1B train-step audio frames/sec"
UNIT = "frames/s"

# algorithmic FLOPs (SURVEY.md §8d / BASELINE.md §3)
BACKBONE_GEMM_FLOPS_PER_FRAME = 1_946_157_056
C0_HEAD_FLOPS_PER_FRAME = 8_400_896
DECODER_FLOPS_PER_SELECTED_FRAME = 7_386_359_808


def attn_flops_per_frame(S):
    return 16 * 2 * S * 2048


def fwd_flops_per_frame(S, fraction=1 / 16):
    return BACKBONE_GEMM_FLOPS_PER_FRAME + attn_flops_per_frame(S) + C0_HEAD_FLOPS_PER_FRAME + \
        DECODER_FLOPS_PER_SELECTED_FRAME * fraction


WORKLOADS = {
    # name: (description, mode, lora_r, targets, default B, default S)
    "c2": ("CSM-1B LoRA r=8 q_proj/v_proj bf16, seq 2048 frames, decoder on 1/16 frames", "lora", 8, None, 2, 2048),
    "c3": ("CSM-1B full fine-tune bf16, decoder 1/16 frame amortisation, data-parallel", "full", 0, None, 2, 2048),
    # BASELINE configs[3]: 4 speakers in every batch — shared backbone adapter, one decoder adapter per speaker, all in
    # one base GEMM per projection (multi-adapter batching, csm/training/multi_speaker_lora.py)
    "c4": ("CSM-1B multi-speaker LoRA r=16 all attn+MLP projections (4 speakers per batch, shared backbone adapter, "
           "per-speaker decoder adapters), long-context 4096 frames", "multi", 16,
           ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"], 16, 4096),
}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.thread, self.index = [], None, None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step_rate(S, steps, warmup, lora_r=8, targets=None, mode="lora"):
    """Times the oracle restatement of the reference path (reference Model/compute_loss arithmetic + decoder term)
    on the host cores with stock PyTorch CPU ops, fp32, all threads: fwd + bwd + AdamW, B=1."""
    from oracle import csm_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.cfg_csm_1b(max(2048, S))
    t0 = time.time()
    model = O.OracleModel(cfg)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(".scale"):
                p.fill_(1.0)
            else:
                p.normal_(0.0, 0.02)
    if mode in ("lora", "multi"):
        O.apply_lora(model, r=lora_r, alpha=16.0, target_modules=targets, seed=1)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.01)
    build_s = time.time() - t0
    times = []
    for i in range(warmup + steps):
        b = O.synthetic_batch(cfg, 1, S, seed=1234 + i)
        t = time.time()
        loss, _ = O.oracle_forward(model, b["input_tokens"], b["input_masks"], b["target_audio_tokens"],
                                   b["frame_idx"])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        if i >= warmup:
            times.append(time.time() - t)
    dt = sum(times) / len(times)
    return {"value": S / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle port of the reference step (fp32, stock PyTorch CPU ops), CSM-1B {mode}, B=1, S={S}, "
                      f"{warmup} warm-up + {steps} timed steps, {dt:.2f} s/step, model build {build_s:.0f} s",
            "ms_per_step": dt * 1e3}


def gpu_stock_step_rate(device, mode, lora_r, targets, B, S, steps=5, warmup=2):
    """SURVEY §8(d) / §0.1 "same box" comparator: the SAME step through stock PyTorch on the SAME GPU — the oracle model
    (reference Model arithmetic + teacher-forced decoder term) in bf16 with cuBLAS GEMMs, SDPA attention,
    F.cross_entropy and torch.optim.AdamW(fused) + clip_grad_norm_, same B / S / frame_idx, CUDA-event timed.  The
    reference has no GPU kernels of its own; this is what its code does when moved to the device unchanged.  Runs
    outside the repo's timed region; nothing of libcsm_b200 is on this path."""
    from oracle import csm_oracle as O
    cfg = O.cfg_csm_1b(max(2048, S))
    with torch.device(device):
        model = O.OracleModel(cfg)
    g = torch.Generator(device=device).manual_seed(0)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(".scale"):
                p.fill_(1.0)
            else:
                p.normal_(0.0, 0.02, generator=g)
    model = model.to(torch.bfloat16)
    if mode in ("lora", "multi"):                  # (stock torch has no multi-adapter batching: one adapter set)
        O.apply_lora(model, r=lora_r, alpha=16.0, target_modules=targets, seed=1)
        model = model.to(device)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.01, fused=True)
    batches = [{k: v.to(device) for k, v in O.synthetic_batch(cfg, B, S, seed=1234 + 97 * i).items()} for i in range(2)]

    def one(i):
        b = batches[i % 2]
        loss, _ = O.oracle_forward(model, b["input_tokens"], b["input_masks"], b["target_audio_tokens"], b["frame_idx"])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
    for i in range(warmup):
        one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        one(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model, opt, params, batches
    torch.cuda.empty_cache()
    return {"value": B * S / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "what": "stock PyTorch on the same GPU: oracle model in bf16 (cuBLAS + SDPA + F.cross_entropy + "
                    "clip_grad_norm_ + torch.optim.AdamW(fused)), eager, same B / S / frame_idx"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    desc, mode, r, targets, _, S = WORKLOADS[args.config]
    S_cpu = args.cpu_seq
    # K and W as asked (each step is the bounded B=1, S=cpu_seq sample: ~0.7 s on the GPU box's host cores, ~2 s on 8
    # cores); capped so that a very long request still ends within a few minutes
    steps, warmup = max(1, min(args.steps, 30)), max(0, min(args.warmup, 5))
    res = cpu_reference_step_rate(S_cpu, steps, warmup, r, targets, mode)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": 1, "seq_len": S_cpu, "parallelism": "cpu",
                       "note": "reference path has no GPU kernels of its own; timed on the host cores "
                               "(bounded sample: B=1, shorter sequence)"},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- B200 arm
def build_model(workload, device, max_seq):
    from csm.models import lora as plora
    from csm.models.model import Model, ModelArgs
    desc, mode, r, targets, _, _ = WORKLOADS[workload]
    with torch.device(device):
        model = Model(ModelArgs("llama-1B", "llama-100M", 128256, 2051, 32))
    model = model.to(torch.bfloat16)
    g = torch.Generator(device=device).manual_seed(0)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(".scale"):
                p.fill_(1.0)
            else:
                p.normal_(0.0, 0.02, generator=g)
    if max_seq > 2048:
        model.backbone.set_max_seq_len(max_seq)
    if mode == "lora":
        plora.apply_lora(model, r=r, alpha=16.0, target_modules=targets, seed=1)
    return model                                   # ("multi": the multi-speaker trainer adds its adapters itself)


# ----------------------------------------------------------------------------- HBM-bound kernels beside the step
def _timed_us(fn, flush, iters=10, warm=3):
    """Median CUDA-event time of fn() with a cold L2.  Before every timed call two READ passes over a 512 MB buffer run
    on the stream: they evict L2 with clean lines (a flush by writing leaves 126 MB of dirty lines whose write-back then
    competes with the timed op's reads) and keep the GPU busy for ~200 us, so the host has enqueued every kernel of the
    op before the first one starts — the op's kernels then run back to back as they do in the graph-replayed step
    (issued against an idle GPU, a 2-launch op of ~40 us is timed with ~10 us of host launch latency inside it)."""
    for _ in range(warm):
        fn()
    words = flush.view(torch.int32)
    ts = []
    for _ in range(iters):
        words.max()
        words.max()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def hbm_kernel_extras(device, peaks):
    """BASELINE config 5 (fused audio_head cross-entropy sweep: 31 codebooks x vocab 2051, D = 1024, N_sel = 64..8192;
    forward = partials + combine, end to end) and K1 (33-way embedding gather-sum), each against both rooflines:
    algorithmic bytes / time vs the measured HBM peak and algorithmic FLOPs / time vs the sustained bf16 peak."""
    from csm import ops
    hbm = peaks.get("hbm_gbs") or 6500.0
    tf = peaks.get("bf16_tflops_sustained") or 1400.0
    flush = torch.zeros(512 << 20, dtype=torch.uint8, device=device)
    g = torch.Generator(device=device).manual_seed(5)
    C, Dd, V, D = 32, 1024, 2051, 2048
    head_t = (torch.randn(C - 1, V, Dd, device=device, generator=g) * 0.02).to(torch.bfloat16)
    c5 = []
    for n_sel in (64, 128, 232, 256, 512, 1024, 2048, 4096, 8192):
        y = torch.randn(n_sel, C, Dd, device=device, generator=g).to(torch.bfloat16)
        codes = torch.randint(0, V, (n_sel, C), device=device, generator=g)
        hv = y[:, 1:]
        fwd = lambda: ops.linear_ce_fwd(hv, head_t, codes[:, 1:], groups=C - 1, tgt_row_stride=C,  # noqa: E731
                                        tgt_group_stride=1)
        us = _timed_us(fwd, flush)
        nbytes = (C - 1) * (Dd * V * 2 + n_sel * (Dd * 2 + 8 + 4))          # weights once + rows
        flops = 2.0 * (C - 1) * n_sel * Dd * V
        row = {"n_sel": n_sel, "fwd_us": us, "fwd_gbs": nbytes / us * 1e-3, "fwd_tflops": flops / us * 1e-6,
               "fwd_frac_hbm": nbytes / us * 1e-3 / hbm, "fwd_frac_tensor": flops / us * 1e-6 / tf}
        _, lse = fwd()
        dy = torch.zeros_like(y)
        bwd = lambda: ops.linear_ce_bwd(hv, head_t, codes[:, 1:], lse, 1.0 / (n_sel * (C - 1)), dh=dy[:, 1:],  # noqa: E731
                                        groups=C - 1, tgt_row_stride=C, tgt_group_stride=1)
        us_b = _timed_us(bwd, flush, iters=6)
        nbytes_b = (C - 1) * (2 * Dd * V * 2 + n_sel * (2 * Dd * 2 + 8 + 4))  # weights twice (logits, dH) + rows + dH
        row.update({"bwd_dh_us": us_b, "bwd_dh_gbs": nbytes_b / us_b * 1e-3,
                    "bwd_dh_tflops": 2 * flops / us_b * 1e-6, "bwd_dh_frac_hbm": nbytes_b / us_b * 1e-3 / hbm,
                    "bwd_dh_frac_tensor": 2 * flops / us_b * 1e-6 / tf})
        c5.append(row)
        del y, codes, dy
    # K1: 4096 audio frames (B=2, S=2048 all-audio) of the CSM-1B tables
    from csm.models.model import Model  # noqa: F401  (tables only)
    audio = (torch.randn(32 * V, D, device=device, generator=g) * 0.02).to(torch.bfloat16)
    text = (torch.randn(128256, D, device=device, generator=g) * 0.02).to(torch.bfloat16)
    k1 = []
    for frames in (4096, 16384):
        tok = torch.zeros(1, frames, 33, dtype=torch.int64, device=device)
        tok[:, :, :32] = torch.randint(0, V, (1, frames, 32), device=device, generator=g)
        msk = torch.zeros(1, frames, 33, dtype=torch.bool, device=device)
        msk[:, :, :32] = True
        us = _timed_us(lambda: ops.embed_gather_sum(tok, msk, audio, text), flush)
        nbytes = frames * (32 * D * 2 + D * 2 + 33 * 8 + 33)                 # 135 465 B / audio frame
        k1.append({"audio_frames": frames, "us": us, "gbs": nbytes / us * 1e-3, "frac_hbm": nbytes / us * 1e-3 / hbm})
    return {"c5_fused_audio_head_ce": c5, "k1_embed_gather_sum": k1, "hbm_peak_gbs": hbm, "tensor_peak_tflops": tf,
            "timing": "CUDA events around the op, median of 10 (6 for backward); before every timed call two read passes "
                      "over a 512 MB buffer evict L2 (clean lines) and let the host enqueue the op's kernels ahead"}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except OSError:
        return {}


def measured_traffic(key):
    """DRAM bytes per launch of the dominant kernel from this round's `ncu --set full` capture (profiles/ncu_traffic.json,
    written by tools/ncu_traffic.py from the .ncu-rep): dram__bytes_read.sum + dram__bytes_write.sum."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        e = t.get(key)
        return (e["dram_bytes"], e.get("source")) if e else (None, None)
    except (OSError, ValueError, KeyError):
        return None, None


# ----------------------------------------------------------------------------- one workload, N ranks
def run_workload(name, args, rank, world, local, device, headline):
    """Builds the model + trainer of workload `name`, warms up, and times K steps device-resident and end to end.
    `headline`: also the GEMM roofline leg.  Returns the record (rank 0) or None."""
    from csm import _lib, ops
    from csm.data.synthetic import synthetic_batch
    from csm.training.lora_trainer import CSMLoRATrainer
    from csm.training.trainer import CSMTrainer
    import torch.distributed as dist
    import logging

    desc, mode, r, targets, B0, S0 = WORKLOADS[name]
    B = (args.batch if headline else None) or B0
    S = (args.seq if headline else None) or S0
    model = build_model(name, device, S)
    outdir = os.path.join("/tmp", f"csm_bench_{os.getpid()}_{name}")
    n_speakers = 4
    if mode == "lora":
        trainer = CSMLoRATrainer("", outdir, lora_r=r, target_modules=targets, model=None, device=str(device))
        trainer.logger.setLevel(logging.ERROR)
        trainer.model = model                      # adapters were applied by build_model (seeded)
        trainer.prepare_optimizer()
        step = lambda batch: trainer.train_step(batch)                      # noqa: E731
        eager_step = lambda batch: trainer._step_impl(batch, 1.0)           # noqa: E731
    elif mode == "multi":
        from csm.training.multi_speaker_lora import MultiSpeakerLoRATrainer
        logging.getLogger("multi_speaker_lora_trainer").setLevel(logging.ERROR)
        logging.getLogger("csm_lora_trainer").setLevel(logging.ERROR)
        multi = MultiSpeakerLoRATrainer("", outdir, speaker_ids=list(range(n_speakers)), lora_r=r,
                                        target_modules=targets, share_backbone=True, share_decoder=False, model=model,
                                        device=str(device))
        multi.logger.setLevel(logging.ERROR)
        trainer = multi.engine
        trainer.logger.setLevel(logging.ERROR)
        multi.prepare_optimizers()
        step = lambda batch: multi.train_step(batch)                        # noqa: E731
        eager_step = lambda batch: trainer._step_impl(batch, 1.0)           # noqa: E731
    else:
        trainer = CSMTrainer("", outdir, device=str(device))
        trainer.logger.setLevel(logging.ERROR)
        trainer.model = model
        trainer.prepare_optimizer()
        step = lambda batch: trainer.train_step(batch)                      # noqa: E731

        def eager_step(batch):
            loss = trainer.train_micro_batch(batch, 1)
            trainer.optimizer_step(1.0)
            return loss
    # A captured step keeps its activations in the graph's private pool NEXT to the eager warm-up's: at c4's batch 16 x
    # 4096 frames (138 GB of activations per step) that does not fit 180 GB, so that workload runs its launches eagerly
    # (kernels of 65 536 rows: launch cost is invisible)
    use_graph = not args.no_graph and B * S <= 32768
    if use_graph:
        trainer.enable_cuda_graph(warmup=args.warmup)

    n_batches = 4
    host = [synthetic_batch(128256, 2051, 32, B, S, seed=1234 + rank + 97 * i) for i in range(n_batches)]
    if mode == "multi":
        for i, b in enumerate(host):               # every batch mixes the speakers
            b["speaker_ids"] = (torch.arange(B) + i) % n_speakers
    # inputs in the data pipeline's compact device format (csm/data/frames.py::pack_tokens: int32 pre-offset table rows
    # + one mask word per frame; bit-identical results, 140 instead of 297 bytes per frame over PCIe)
    from csm.data.frames import compact_batch
    host = [compact_batch(b, 2051, pin=False) for b in host]
    host = [{k: v.pin_memory() for k, v in b.items()} for b in host]
    resident = [{k: v.to(device) for k, v in b.items()} for b in host]
    n_sel = int(host[0]["frame_idx"].shape[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def phase(what):
        if os.environ.get("CSM_BENCH_TRACE"):
            print(f"[bench rank {rank}] {name}: {what}", file=sys.stderr, flush=True)

    def timed(fn, K):
        """-> (max over ranks of the device time of K calls, [per-rank times])"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        mine = torch.tensor([e0.elapsed_time(e1)], device=device)
        every = [mine]
        if world > 1:
            every = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(every, mine)
        barrier()
        per_rank = [float(t.item()) for t in every]
        return max(per_rank), per_rank

    n_warm = args.warmup + (2 if use_graph else 0)               # graph mode: W eager steps, then capture + 1 replay
    for i in range(n_warm):
        phase(f"warm-up step {i}")
        step(resident[i % n_batches])
        if os.environ.get("CSM_BENCH_TRACE"):
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    phase("timed region")

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = _lib.launch_count()
    ms_total, per_rank = timed(lambda i: step(resident[i % n_batches]), args.steps)
    launches = _lib.launch_count() - l0
    graphed = getattr(trainer, "_graphed", None)
    if use_graph and graphed is not None and graphed.graph is not None:
        launches = graphed.kernels_per_replay * args.steps   # replayed graph nodes (counted at capture)
    ms_step = ms_total / args.steps
    frames_per_step = world * B * S
    value = frames_per_step / (ms_step * 1e-3)

    # end to end through the public trainer API with HOST batches (H2D of inputs + D2H of the loss each step)
    e2e = None
    if not args.no_e2e:
        def e2e_step(i):
            loss = step(host[i % n_batches])
            return float(loss)                     # D2H read of the step's result
        for i in range(2):
            e2e_step(i)
        ms_e2e = timed(e2e_step, args.steps)[0] / args.steps
        h2d = sum(v.numel() * v.element_size() for v in host[0].values())
        e2e = {"value": frames_per_step / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e}
    clk = clocks.stop() if rank == 0 else None        # sampled across both timed regions (device-resident and e2e)

    # exposed part of the gradient exchange: one eager step with CUDA events around GradSynchronizer.finish() (the
    # main stream waits there for whatever part of the collectives the backward did not hide) and per-bucket events
    exchange = None
    if world > 1:
        sync = trainer._sync
        sync.trace = []
        eager_step(resident[0])
        torch.cuda.synchronize()
        tr, sync.trace = sync.trace, None
        fin = [t for t in tr if t[0] == "finish"]
        exposed = fin[-1][1].elapsed_time(fin[-1][2]) if fin else None
        buckets = [{"bucket": t[1], "mbytes": t[2] / 1e6, "allreduce_ms": t[3].elapsed_time(t[4])}
                   for t in tr if t[0] == "bucket"]
        mine = torch.tensor([exposed or 0.0], device=device)
        dist.all_reduce(mine, op=dist.ReduceOp.MAX)
        exchange = {"exposed_ms_max_over_ranks": float(mine.item()), "exposed_ms_rank0": exposed,
                    "mode": "bucketed, overlapped with backward" if sync.bucketed else "one flat all-reduce after backward",
                    "buckets_rank0": buckets[:40],
                    "note": "eager step (not the graph replay); exposed = device time of GradSynchronizer.finish() on "
                            "the main stream; allreduce_ms = duration of each bucket's NCCL kernel on the side stream"}

    # roofline of the dominant kernel (tcgen05 GEMM): CUDA events around every GEMM launch of three more steps, the one
    # with the least total GEMM time is reported (power-capped GPU: the eager cadence lets the clocks wander ~8 %)
    roof = None
    if headline:
        prof = []
        for i in range(3):
            if rank == 0:
                ops.gemm_profile_start()
            eager_step(resident[i % n_batches])
            torch.cuda.synchronize()
            if rank == 0:
                prof.append(ops.gemm_profile_stop())
        if rank == 0:
            flops, gms, n_g = min(prof, key=lambda q: q[1])
            peaks = load_peaks()
            peak = peaks.get("bf16_tflops_sustained")
            src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
            if peak is None:
                peak, src = 1400.0, "fallback (B200_PROFILING.md sustained)"
            achieved = flops / (gms * 1e-3) / 1e12 if gms > 0 else 0.0
            step_flops = fwd_flops_per_frame(S) * (3 if mode == "full" else 2) * B * S
            key = f"gemm_gateup_{B * S}x16384x2048"
            traffic, tsrc = measured_traffic(key)
            roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 cta_group::2 / TMEM / TMA bf16 GEMM)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic,
                    "traffic_note": f"dram bytes per launch of the {B * S}x16384x2048 fused gate/up GEMM, the largest "
                                    f"instantiation ({tsrc or 'no ncu capture for this shape: null'})",
                    "peak_source": src, "gemm_launches_per_step": n_g, "gemm_ms_per_step": gms,
                    "gemm_share_of_step": gms / ms_step,
                    "step_model_tflops": step_flops / (ms_step * 1e-3) / 1e12,
                    "step_frac_of_peak": step_flops / (ms_step * 1e-3) / 1e12 / peak}

    rec = None
    if rank == 0:
        spread = {"min": min(per_rank) / args.steps, "max": max(per_rank) / args.steps,
                  "mean": sum(per_rank) / len(per_rank) / args.steps,
                  "per_rank": [t / args.steps for t in per_rank]}
        rec = {"value": value, "ms_per_step": ms_step, "name": name, "workload": desc, "mode": mode, "batch_per_gpu": B,
               "max_memory_gb": torch.cuda.max_memory_allocated(device) / 1e9, "cuda_graph": use_graph,
               "seq_len": S, "decoder_frames_per_gpu": n_sel, "e2e": e2e, "gpu_launches": launches, "clocks": clk,
               "rank_ms_per_step": spread, "exchange": exchange, "roofline": roof}
    # free everything this workload holds on the device before the next one is built
    del trainer, model, resident, host, step, eager_step
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return rec, (B, S, mode, r, targets)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU")
    ap.add_argument("--seq", type=int, default=None)
    ap.add_argument("--cpu-seq", type=int, default=256, help="sequence length of the bounded CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stock-baseline", action="store_true", help="skip the stock-PyTorch-on-GPU comparator")
    ap.add_argument("--no-fullft", action="store_true", help="skip the c3 (full fine-tune, data-parallel) sub-record")
    ap.add_argument("--no-extras", action="store_true", help="skip the c5 fused-CE sweep / K1 gather GB/s block")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel launch from Python (no CUDA graph)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    from csm import _lib
    from csm.training import dp
    import torch.distributed as dist

    rank, world, local = dp.init_distributed()
    if world != args.gpus and rank == 0 and world > 1:
        print(f"warning: WORLD_SIZE={world} but --gpus {args.gpus}", file=sys.stderr)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    lib = _lib.load()
    if lib.csm_device_supported() != 1:
        raise RuntimeError("bench.py needs an sm_100 (B200) device: libcsm_b200 has no fallback path")

    main_rec, (B, S, mode, r, targets) = run_workload(args.config, args, rank, world, local, device, headline=True)

    # BASELINE configs[2]: CSM-1B full fine-tune, data-parallel — the configuration multi-GPU scaling is quoted on.
    # Measured in the same run at every N (N = 1 gives the denominator of the scaling efficiency).
    fullft = None
    if not args.no_fullft and args.config != "c3":
        fullft, _ = run_workload("c3", args, rank, world, local, device, headline=False)

    stock = cpu = extras = None
    if rank == 0 and world == 1:
        # stock PyTorch on the same GPU, outside every timed region of this repo's path
        if not args.no_stock_baseline:
            try:
                stock = gpu_stock_step_rate(device, mode, r or 8, targets, B, S)
                stock["speedup_of_this_repo"] = main_rec["value"] / stock["value"]
            except torch.OutOfMemoryError as e:      # the unfused path materialises [B,S,33,D] and full logits
                stock = {"unavailable": f"out of memory: {str(e)[:120]}"}
                torch.cuda.empty_cache()
        if not args.no_extras:
            extras = hbm_kernel_extras(device, load_peaks())
        # CPU baseline: bounded sample of the same workload on the host cores
        if not args.no_cpu_baseline:
            torch.cuda.empty_cache()
            res = cpu_reference_step_rate(args.cpu_seq, 2, 1, r or 8, targets, mode)
            cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
            cpu["note"] = ("the sample runs S=%d where causal attention is %.0fx cheaper per frame than at S=%d; "
                           "attention is ~5 %% of the step's FLOPs, so frames/s at S=%d would be ~%.0f %% lower"
                           % (args.cpu_seq, S / args.cpu_seq, S, S,
                              100 * (1 - fwd_flops_per_frame(args.cpu_seq) / fwd_flops_per_frame(S))))

    if rank == 0:
        m = main_rec
        line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": m["workload"], "name": m["name"], "batch_per_gpu": m["batch_per_gpu"],
                           "global_batch": m["batch_per_gpu"] * world, "seq_len": m["seq_len"],
                           "decoder_frames_per_gpu": m["decoder_frames_per_gpu"],
                           "parallelism": f"dp{world}" if world > 1 else "single",
                           "cuda_graph": m["cuda_graph"],
                           "input_format": "compact (int32 pre-offset table rows [B,S,33] + one int64 mask word per "
                                           "frame; csm/data/frames.py::pack_tokens), targets int64",
                           "precision": "bf16 GEMM / attention operands, fp32 accumulation, fp32 residual stream, "
                                        "fp32 master weights + fp32 AdamW moments",
                           "l2": "per-step working set (3.1 GB weights + >4 GB activations) exceeds the 126 MB L2; "
                                 "4 distinct input batches cycled"},
                "clocks": m["clocks"], "e2e": m["e2e"], "gpu_launches": m["gpu_launches"],
                "gpu_launches_per_step": m["gpu_launches"] / args.steps, "rank_ms_per_step": m["rank_ms_per_step"],
                "exchange": m["exchange"], "max_memory_gb": m["max_memory_gb"], "roofline": m["roofline"],
                "cpu_baseline": cpu,
                "gpu_stock_baseline": stock, "fullft": fullft, "extra": extras}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)       # skip NCCL/graph teardown (destroy_process_group can hang while captured graphs hold NCCL work)


if __name__ == "__main__":
    main()
