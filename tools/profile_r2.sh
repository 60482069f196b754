#!/usr/bin/env bash
# Round-2 profiling pass (run under gpurun, ONE GPU): launch list of an eager c2 step + `ncu --set full` captures of the
# dominant kernels.  Each target runs plainly first (exit code checked) and only then under ncu.
set -u
OUT=gpurun_out/r2p
mkdir -p $OUT
STEP="python bench.py --no-graph --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-stock-baseline --no-extras --no-fullft"
$STEP > $OUT/plain_step.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $OUT/launches_c2_step.csv $STEP > $OUT/ncu_step.log 2>&1
echo "launch list rc=$?"
for t in swiglu attn ce embed; do
  case $t in
    swiglu) pat="gemm_tc_kernel";;
    attn) pat="attn_";;
    ce) pat="gemm_tc_kernel|ce_combine";;
    embed) pat="embed_gather";;
  esac
  python tools/ncu_target.py $t > $OUT/plain_$t.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k "regex:$pat" -c 8 -o $OUT/$t -f python tools/ncu_target.py $t > $OUT/ncu_$t.log 2>&1
  echo "$t rc=$?"
done
ls -la $OUT
