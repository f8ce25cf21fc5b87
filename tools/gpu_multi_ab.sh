# N-GPU bench under different NCCL CTA limits: bash tools/gpu_multi_ab.sh N "val1 val2 ..."   ("-" = NCCL default)
N=${1:-2}
for v in $2; do
  if [ "$v" = "-" ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$v; fi
  timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/multi_err.log | tail -1 | python -c "
import json,sys
t=sys.stdin.read().strip()
try:
    d=json.loads(t); print('N=$N NCCL_MAX_CTAS=$v ->', round(d['value']), 'samples/s', round(d['ms_per_step'],4), 'ms/step')
    open('gpurun_out/multi_ab.jsonl','a').write(t+'\n')
except Exception as e: print('FAILED', e, t[:200])
"
done
