#!/usr/bin/env bash
# GPU call 5: fused-CE load batching A/B (kernel durations per launch), CE tests with the batch on
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
CSM_CE_LOAD_BATCH=2 timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "linear_ce" > gpurun_out/c5_tests_ce.log 2>&1
echo "ce tests rc=$?" | tee gpurun_out/c5_status.txt
for lb in 2 1; do
  CSM_CE_LOAD_BATCH=$lb python tools/ce_sweep_target.py > gpurun_out/c5_ce_plain_$lb.log 2>&1 && \
  CSM_CE_LOAD_BATCH=$lb ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none --csv \
      --log-file gpurun_out/c5_ce_launches_lb$lb.csv python tools/ce_sweep_target.py > gpurun_out/c5_ce_ncu_$lb.log 2>&1
  echo "ce list lb=$lb rc=$?" | tee -a gpurun_out/c5_status.txt
done
