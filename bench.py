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
    "c4": ("CSM-1B multi-speaker LoRA r=16 all attn+MLP projections, long-context 4096 frames", "lora", 16,
           ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"], 2, 4096),
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
    if mode == "lora":
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
    if mode == "lora":
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
    return model


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
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel launch from Python (no CUDA graph)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    from csm import _lib, ops
    from csm.data.synthetic import synthetic_batch
    from csm.training import dp
    from csm.training.lora_trainer import CSMLoRATrainer
    from csm.training.trainer import CSMTrainer
    import torch.distributed as dist

    rank, world, local = dp.init_distributed()
    if world != args.gpus and rank == 0 and world > 1:
        print(f"warning: WORLD_SIZE={world} but --gpus {args.gpus}", file=sys.stderr)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    lib = _lib.load()
    if lib.csm_device_supported() != 1:
        raise RuntimeError("bench.py needs an sm_100 (B200) device: libcsm_b200 has no fallback path")

    desc, mode, r, targets, B0, S0 = WORKLOADS[args.config]
    B, S = args.batch or B0, args.seq or S0
    model = build_model(args.config, device, S)
    outdir = os.path.join("/tmp", f"csm_bench_{os.getpid()}")
    import logging
    if mode == "lora":
        trainer = CSMLoRATrainer("", outdir, lora_r=r, target_modules=targets, model=None, device=str(device))
        trainer.logger.setLevel(logging.ERROR)
        trainer.model = model                      # adapters were applied by build_model (seeded)
        trainer.prepare_optimizer()
        step = lambda batch: trainer.train_step(batch)                      # noqa: E731
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
    if not args.no_graph:
        trainer.enable_cuda_graph(warmup=args.warmup)

    n_batches = 4
    host = [synthetic_batch(128256, 2051, 32, B, S, seed=1234 + rank + 97 * i) for i in range(n_batches)]
    host = [{k: v.pin_memory() for k, v in b.items()} for b in host]
    resident = [{k: v.to(device) for k, v in b.items()} for b in host]
    n_sel = int(host[0]["frame_idx"].shape[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def phase(name):
        if os.environ.get("CSM_BENCH_TRACE"):
            print(f"[bench rank {rank}] {name}", file=sys.stderr, flush=True)

    def timed(fn, K):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    # ---- warm-up (also instantiates optimizer state, cuda modules)
    n_warm = args.warmup + (2 if not args.no_graph else 0)       # graph mode: W eager steps, then capture + 1 replay
    for i in range(n_warm):
        phase(f"warm-up step {i}")
        step(resident[i % n_batches])
        if os.environ.get("CSM_BENCH_TRACE"):
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    phase("timed region")

    # ---- device-resident throughput (value)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = _lib.launch_count()
    ms_total = timed(lambda i: step(resident[i % n_batches]), args.steps)
    launches = _lib.launch_count() - l0
    if not args.no_graph and getattr(trainer, "_graphed", None) is not None and trainer._graphed.graph is not None:
        launches = trainer._graphed.kernels_per_replay * args.steps   # replayed graph nodes (counted at capture)
    ms_step = ms_total / args.steps
    frames_per_step = world * B * S
    value = frames_per_step / (ms_step * 1e-3)

    # ---- end to end through the public trainer API with HOST batches (H2D of inputs + D2H of the loss each step)
    e2e = None
    if not args.no_e2e:
        def e2e_step(i):
            loss = step(host[i % n_batches])
            return float(loss)                     # D2H read of the step's result
        for i in range(2):
            e2e_step(i)
        ms_e2e = timed(e2e_step, args.steps) / args.steps
        h2d = sum(v.numel() * v.element_size() for v in host[0].values())
        e2e = {"value": frames_per_step / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e}

    clk = clocks.stop() if rank == 0 else None        # sampled across both timed regions (device-resident and e2e)

    # ---- roofline of the dominant kernel (tcgen05 GEMM): CUDA events around every GEMM launch of two more steps
    # (every rank runs the two steps — they contain the gradient all-reduce — only rank 0 instruments them)
    roof = None
    # three instrumented eager steps; the one with the least total GEMM time is reported (the GPU is power-capped and
    # the eager launch cadence lets the clocks wander between steps: the spread is ~8 %)
    prof = []
    for i in range(3):
        if rank == 0:
            ops.gemm_profile_start()
        eager_step(resident[i % n_batches])
        torch.cuda.synchronize()
        if rank == 0:
            prof.append(ops.gemm_profile_stop())
    if rank == 0:
        flops, gms, n_g = min(prof, key=lambda r: r[1])
        flops, gms, n_g = 2 * flops, 2 * gms, 2 * n_g          # (the fields below are per two steps)
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        peak = peaks.get("bf16_tflops_sustained")
        src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
        if peak is None:
            peak, src = 1400.0, "fallback (B200_PROFILING.md sustained)"
        achieved = flops / (gms * 1e-3) / 1e12 if gms > 0 else 0.0
        step_flops = fwd_flops_per_frame(S) * (2 if mode == "lora" else 3) * B * S
        # DRAM bytes of ONE launch of the dominant instantiation (CTA-pair kernel on the fused gate/up forward GEMM,
        # 4096 x 16384 x 2048) from the committed `ncu --set full` capture: dram__bytes_read.sum + dram__bytes_write.sum
        # = 84.3 MB + 90.8 MB (profiles/r1_ncu_gemm_cta_pair.txt); algorithmic operand bytes of that launch: 218.1 MB
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 cta_group::2 / TMEM / TMA bf16 GEMM)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": 175.2e6 if (S == 2048 and B == 2) else None,
                "traffic_note": "per launch of the 4096x16384x2048 gate/up GEMM (ncu, profiles/r1_ncu_gemm_cta_pair.txt)",
                "peak_source": src,
                "gemm_launches_per_step": n_g // 2, "gemm_ms_per_step": gms / 2, "gemm_share_of_step": (gms / 2) / ms_step,
                "step_model_tflops": step_flops / (ms_step * 1e-3) / 1e12,
                "step_frac_of_peak": step_flops / (ms_step * 1e-3) / 1e12 / peak}

    # ---- stock PyTorch on the same GPU (rank 0, N=1 only), outside every timed region of this repo's path
    stock = None
    if rank == 0 and world == 1 and not args.no_stock_baseline:
        try:
            stock = gpu_stock_step_rate(device, mode, r or 8, targets, B, S)
            stock["speedup_of_this_repo"] = value / stock["value"]
        except torch.OutOfMemoryError as e:      # the unfused path materialises [B,S,33,D] and full logits
            stock = {"unavailable": f"out of memory: {str(e)[:120]}"}
            torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del resident
        torch.cuda.empty_cache()
        res = cpu_reference_step_rate(args.cpu_seq, 2, 1, r or 8, targets, mode)
        cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": desc, "name": args.config, "batch_per_gpu": B, "global_batch": B * world,
                           "seq_len": S, "decoder_frames_per_gpu": n_sel,
                           "parallelism": f"dp{world}" if world > 1 else "single",
                           "cuda_graph": not args.no_graph,
                           "l2": "per-step working set (3.1 GB weights + >4 GB activations) exceeds the 126 MB L2; "
                                 "4 distinct input batches cycled"},
                "clocks": clk, "e2e": e2e, "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
                "roofline": roof, "cpu_baseline": cpu, "gpu_stock_baseline": stock}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)       # skip NCCL/graph teardown (destroy_process_group can hang while captured graphs hold NCCL work)


if __name__ == "__main__":
    main()
