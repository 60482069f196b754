#!/usr/bin/env bash
# GPU call 14: experiments compiled out of the GEMM kernels — GEMM / CE / parity tests, the experiments build's own
# tests, step time
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/c14_tests_gpu.log 2>&1
echo "pytest -m gpu rc=$?" | tee gpurun_out/c14_status.txt
tail -2 gpurun_out/c14_tests_gpu.log | tee -a gpurun_out/c14_status.txt
CSM_B200_LIB=$PWD/csm-train-pytorch_b200/libcsm_b200_exp.so CSM_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -k "stream_k or narrow_tail" > gpurun_out/c14_tests_exp.log 2>&1
echo "experiments build tests rc=$?" | tee -a gpurun_out/c14_status.txt
tail -2 gpurun_out/c14_tests_exp.log | tee -a gpurun_out/c14_status.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-stock-baseline --no-extras --no-e2e > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err
echo "bench rc=$?" | tee -a gpurun_out/c14_status.txt
grep -o '"ms_per_step": [0-9.]*' gpurun_out/c14_bench.json | head -3 | tee -a gpurun_out/c14_status.txt
