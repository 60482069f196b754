mkdir -p gpurun_out/r2g
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c3 --no-e2e --steps 15 --warmup 3 > gpurun_out/r2g/c3_$tag.json 2> gpurun_out/r2g/c3_$tag.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2g/c3_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['ms_per_step'],2), 'exposed', round(d['exchange']['exposed_ms_rank0'],2), 'ar ms sum', round(sum(b['allreduce_ms'] for b in d['exchange']['buckets_rank0']),2))
PY
}
run default FOO=1
run ctas4 NCCL_MAX_CTAS=4
run ctas8_res8 NCCL_MAX_CTAS=8 CSM_DP_RESERVED_SMS=8
run ctas16_res16 NCCL_MAX_CTAS=16 CSM_DP_RESERVED_SMS=16
run ctas2 NCCL_MAX_CTAS=2
