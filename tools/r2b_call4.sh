#!/usr/bin/env bash
# GPU call 4: fused-CE weights L2 row prefetch A/B (kernel durations + DRAM bytes per launch), CE tests
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "linear_ce" > gpurun_out/c4_tests_ce.log 2>&1
echo "ce tests rc=$?" | tee gpurun_out/c4_status.txt
for pf in 1 0; do
  CSM_CE_L2_PREFETCH=$pf python tools/ce_sweep_target.py > gpurun_out/c4_ce_plain_$pf.log 2>&1 && \
  CSM_CE_L2_PREFETCH=$pf ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none --csv \
      --log-file gpurun_out/c4_ce_launches_pf$pf.csv python tools/ce_sweep_target.py > gpurun_out/c4_ce_ncu_$pf.log 2>&1
  echo "ce list pf=$pf rc=$?" | tee -a gpurun_out/c4_status.txt
done
