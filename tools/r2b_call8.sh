#!/usr/bin/env bash
# GPU call 8: fused-CE weights: TMA L2 prefetch distance sweep (kernel durations per launch)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for d in 0 4 8 12; do
  CSM_CE_B_PREFETCH=$d ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none --csv \
      --log-file gpurun_out/c8_ce_launches_d$d.csv python tools/ce_sweep_target.py > gpurun_out/c8_ce_ncu_$d.log 2>&1
  echo "dist=$d rc=$?" | tee -a gpurun_out/c8_status.txt
done
CSM_CE_B_PREFETCH=8 timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "linear_ce" > gpurun_out/c8_tests_ce.log 2>&1
echo "ce tests rc=$?" | tee -a gpurun_out/c8_status.txt
