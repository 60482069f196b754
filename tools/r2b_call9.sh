#!/usr/bin/env bash
# GPU call 9: RMSNorm backward with the scale-gradient partials in warp-private smem; CE combine launched with PDL
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "rmsnorm or linear_ce" > gpurun_out/c9_tests.log 2>&1
echo "tests rc=$?" | tee gpurun_out/c9_status.txt
timeout 900 python -m pytest tests/test_parity_csm1b_gpu.py tests/test_model_parity_gpu.py -x -q > gpurun_out/c9_tests_parity.log 2>&1
echo "parity rc=$?" | tee -a gpurun_out/c9_status.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c9_launches_rmsnorm.csv python tools/ncu_target.py rmsnorm > gpurun_out/c9_ncu_rmsnorm.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-stock-baseline --no-e2e > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
echo "bench rc=$?" | tee -a gpurun_out/c9_status.txt
