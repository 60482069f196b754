#!/usr/bin/env bash
# GPU call 7 (2 GPUs): extras timing check at N=1, then the driver-style 2-GPU line (c2 + c3 sub-record)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-stock-baseline --no-fullft --no-e2e > gpurun_out/c7_bench_extras.json 2> gpurun_out/c7_bench_extras.err
echo "extras rc=$?" | tee gpurun_out/c7_status.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c7_bench_dp2.json 2> gpurun_out/c7_bench_dp2.err
echo "dp2 rc=$?" | tee -a gpurun_out/c7_status.txt
