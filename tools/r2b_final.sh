#!/usr/bin/env bash
# final validation of the round: full GPU test suite, smoke, driver-style bench at N=1
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/final_tests_gpu.log 2>&1
echo "pytest -m gpu rc=$?" | tee gpurun_out/final_status.txt
tail -3 gpurun_out/final_tests_gpu.log | tee -a gpurun_out/final_status.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/final_status.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
echo "bench rc=$?" | tee -a gpurun_out/final_status.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
echo "reference arm rc=$?" | tee -a gpurun_out/final_status.txt
