#!/usr/bin/env bash
# GPU call 11 (2 GPUs): compact device token format — full GPU suite, bench at N=1 and N=2
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/c11_tests_gpu.log 2>&1
echo "pytest -m gpu rc=$?" | tee gpurun_out/c11_status.txt
tail -3 gpurun_out/c11_tests_gpu.log | tee -a gpurun_out/c11_status.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-stock-baseline --no-extras > gpurun_out/c11_bench_n1.json 2> gpurun_out/c11_bench_n1.err
echo "bench n1 rc=$?" | tee -a gpurun_out/c11_status.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-stock-baseline --no-extras > gpurun_out/c11_bench_dp2.json 2> gpurun_out/c11_bench_dp2.err
echo "bench dp2 rc=$?" | tee -a gpurun_out/c11_status.txt
