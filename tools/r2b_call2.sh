#!/usr/bin/env bash
# round 2, session 2, GPU call 2: skinny kernels with batched loads, PDL variants (early trigger / no trigger), per-kernel times
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "skinny or rmsnorm or programmatic" > gpurun_out/c2_tests_ops.log 2>&1
echo "ops tests rc=$?" | tee gpurun_out/c2_status.txt
CSM_PDL=0 timeout 300 python tools/bench_skinny.py > gpurun_out/c2_skinny_pdl0.json 2> gpurun_out/c2_skinny.err
echo "bench_skinny rc=$?" | tee -a gpurun_out/c2_status.txt
CSM_PDL=1 timeout 300 python tools/bench_skinny.py > gpurun_out/c2_skinny_pdl1.json 2>> gpurun_out/c2_skinny.err
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stock-baseline --no-fullft --no-extras --no-e2e"
NT=$PWD/csm-train-pytorch_b200/libcsm_b200_nt.so
i=0
for cfg in "0 1 -" "0 0 -" "1 1 nt" "1 1 -" "0 1 -" "1 1 nt"; do
  set -- $cfg
  i=$((i+1))
  if [ "$3" = "nt" ]; then export CSM_B200_LIB=$NT; else unset CSM_B200_LIB; fi
  CSM_PDL=$1 CSM_SKINNY=$2 timeout 300 $B > gpurun_out/c2_bench_${i}_pdl$1_sk$2_$3.json 2>> gpurun_out/c2_bench.err
  echo "bench pdl=$1 skinny=$2 lib=$3 rc=$?" | tee -a gpurun_out/c2_status.txt
done
unset CSM_B200_LIB
for f in gpurun_out/c2_bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -1)"; done | tee -a gpurun_out/c2_status.txt
# launch list (per-kernel durations, serialised) of the skinny / rmsnorm targets
for t in skinny rmsnorm; do
  CSM_PDL=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c2_launches_$t.csv python tools/ncu_target.py $t > gpurun_out/c2_ncu_$t.log 2>&1
done
CSM_PDL=0 CSM_SKINNY=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c2_launches_skinny_off.csv python tools/ncu_target.py skinny > gpurun_out/c2_ncu_skinny_off.log 2>&1
echo "ncu lists done" | tee -a gpurun_out/c2_status.txt
