#!/usr/bin/env bash
# GPU call 16: ncu --set full of the fused-CE forward at N_sel = 64 (single-CTA tiles) and 232 (CTA pairs)
set -u
OUT=gpurun_out/r2b
mkdir -p $OUT
cd "$(dirname "$0")/.."
python tools/ce_sweep_target.py > $OUT/plain_ce_sweep.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k "regex:gemm_tc_kernel|ce_combine" -c 24 -o $OUT/ce_sweep -f python tools/ce_sweep_target.py > $OUT/ncu_ce_sweep.log 2>&1
echo "ce rc=$?"
ncu -i $OUT/ce_sweep.ncu-rep --page raw --csv > $OUT/ce_sweep_raw.csv 2>/dev/null
python tools/ncu_stalls.py $OUT/ce_sweep.ncu-rep "gemm_tc_kernel" 12 > $OUT/ce_sweep_stalls.txt 2>&1
ls -la $OUT | tail -5
