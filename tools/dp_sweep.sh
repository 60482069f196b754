#!/usr/bin/env bash
# A/B sweep of SM reservation for the overlapped gradient all-reduce (full fine-tune, N GPUs):
#   tools/dp_sweep.sh <ngpus> "<reserved> <nccl_max_ctas>" ...
N=$1; shift
port=29600
for cfg in "$@"; do
  set -- $cfg
  port=$((port + 1))
  if [ "$2" != "0" ]; then export NCCL_MAX_CTAS=$2; else unset NCCL_MAX_CTAS; fi
  CSM_DP_RESERVED_SMS=$1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --config c3 --steps 10 --warmup 3 --no-e2e \
    > gpurun_out/dp_sweep_$1_$2.json 2> gpurun_out/dp_sweep_$1_$2.err
  python - "$1" "$2" <<PY
import json, sys
try:
    d = json.loads(open("gpurun_out/dp_sweep_%s_%s.json" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
    print("reserved=%s nccl_max_ctas=%s ms/step=%.2f frames/s=%.0f" % (sys.argv[1], sys.argv[2], d["ms_per_step"], d["value"]))
except Exception as e:
    print("reserved=%s nccl_max_ctas=%s FAILED %r" % (sys.argv[1], sys.argv[2], e))
PY
done
