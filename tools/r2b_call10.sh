#!/usr/bin/env bash
# GPU call 10: dK/dV kernel on a side stream next to the dQ kernel — microbenchmark, tests, step A/B
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for ov in 0 1; do CSM_ATTN_BWD_OVERLAP=$ov python tools/bench_attn_bwd.py > gpurun_out/c10_attn_bwd_$ov.json 2> gpurun_out/c10_attn_bwd_$ov.err; cat gpurun_out/c10_attn_bwd_$ov.json; done | tee gpurun_out/c10_status.txt
CSM_ATTN_BWD_OVERLAP=1 timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_parity_csm1b_gpu.py tests/test_trainers_gpu.py -x -q -k "attention or c2 or c3 or graph or packed" > gpurun_out/c10_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/c10_status.txt
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-baseline --no-extras --no-e2e"
for ov in 0 1 0 1; do
  CSM_ATTN_BWD_OVERLAP=$ov timeout 300 $B > gpurun_out/c10_bench_ov${ov}_$RANDOM.json 2>> gpurun_out/c10_bench.err
done
for f in gpurun_out/c10_bench_ov*.json; do python - "$f" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.strip().startswith('{'):
        d=json.loads(line); print(sys.argv[1], round(d['ms_per_step'],3), 'c3', round(d['fullft']['ms_per_step'],3), d['clocks']['sm_mhz'])
PY
done | tee -a gpurun_out/c10_status.txt
