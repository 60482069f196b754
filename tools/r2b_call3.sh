#!/usr/bin/env bash
# round 2, session 2, GPU call 3: evidence pass — launch list of an eager c2 step, ncu --set full of the new kernels,
# CE forward launch list at the c5 sizes, the bench line with the graph-timed c5 sweep
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bash tools/profile_r2b.sh > gpurun_out/c3_profile.log 2>&1
echo "profile rc=$?" | tee gpurun_out/c3_status.txt
python tools/ce_sweep_target.py > gpurun_out/c3_ce_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/c3_ce_launches.csv python tools/ce_sweep_target.py > gpurun_out/c3_ce_ncu.log 2>&1
echo "ce list rc=$?" | tee -a gpurun_out/c3_status.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c3_bench_n1.json 2> gpurun_out/c3_bench_n1.err
echo "bench rc=$?" | tee -a gpurun_out/c3_status.txt
