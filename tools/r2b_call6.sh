#!/usr/bin/env bash
# GPU call 6: does the 2 KB row pitch of the audio-head weights camp on HBM channels?  CE forward with padded row pitch.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for pad in 0 8 16 64 520; do
  CE_PAD=$pad ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none --csv \
      --log-file gpurun_out/c6_ce_launches_pad$pad.csv python tools/ce_sweep_target.py > gpurun_out/c6_ce_ncu_$pad.log 2>&1
  echo "pad=$pad rc=$?" | tee -a gpurun_out/c6_status.txt
done
