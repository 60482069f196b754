#!/usr/bin/env bash
# round 2, session 2, GPU call 1: new kernels' tests, parity at CSM-1B dimensions, A/B of PDL / skinny kernels
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -k "skinny or rmsnorm or split_reduction or programmatic or attention_fwd_bwd or gemm_cta_pair" > gpurun_out/c1_tests_ops.log 2>&1
echo "ops tests rc=$?" | tee gpurun_out/c1_status.txt
timeout 900 python -m pytest tests/test_model_parity_gpu.py -x -q > gpurun_out/c1_tests_model.log 2>&1
echo "model parity rc=$?" | tee -a gpurun_out/c1_status.txt
timeout 900 python -m pytest tests/test_parity_csm1b_gpu.py -x -q -k "c2" > gpurun_out/c1_tests_csm1b.log 2>&1
echo "csm1b c2 parity rc=$?" | tee -a gpurun_out/c1_status.txt
timeout 300 python tools/bench_skinny.py > gpurun_out/c1_skinny.json 2> gpurun_out/c1_skinny.err
echo "bench_skinny rc=$?" | tee -a gpurun_out/c1_status.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-baseline --no-fullft --no-extras --no-e2e"
for cfg in "0 0" "0 1" "1 1" "2 1" "1 0" "0 0" "2 1"; do
  set -- $cfg
  CSM_PDL=$1 CSM_SKINNY=$2 timeout 300 $B > gpurun_out/c1_bench_pdl$1_sk$2_$RANDOM.json 2>> gpurun_out/c1_bench.err
  echo "bench pdl=$1 skinny=$2 rc=$?" | tee -a gpurun_out/c1_status.txt
done
grep -h -o '"ms_per_step": [0-9.]*' gpurun_out/c1_bench_pdl*.json | head -20
for f in gpurun_out/c1_bench_pdl*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -1)"; done | tee -a gpurun_out/c1_status.txt
