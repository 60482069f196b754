#!/usr/bin/env bash
# Round 2, second session: launch list of an eager c2 step with the streaming LoRA kernels + PDL, and `ncu --set full`
# captures of the new kernels (skinny rowdot / coldot, RMSNorm backward).  Each target runs plainly first.
set -u
OUT=gpurun_out/r2b
mkdir -p $OUT
STEP="python bench.py --no-graph --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-stock-baseline --no-extras --no-fullft"
$STEP > $OUT/plain_step.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $OUT/launches_c2_step.csv $STEP > $OUT/ncu_step.log 2>&1
echo "launch list rc=$?"
python tools/launch_summary.py $OUT/launches_c2_step.csv > $OUT/launches_c2_step_summary.txt 2>&1
for t in skinny rmsnorm; do
  case $t in
    skinny) pat="skinny_";;
    rmsnorm) pat="rmsnorm_";;
  esac
  python tools/ncu_target.py $t > $OUT/plain_$t.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k "regex:$pat" -c 8 -o $OUT/$t -f python tools/ncu_target.py $t > $OUT/ncu_$t.log 2>&1
  echo "$t rc=$?"
  ncu -i $OUT/$t.ncu-rep --page raw --csv > $OUT/${t}_raw.csv 2>/dev/null
done
ls -la $OUT
