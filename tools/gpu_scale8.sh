# 8-GPU scaling lines: peer vs NCCL transport (batch 256/GPU), batch 512/GPU, and the 1-GPU reference points
N=${1:-8}
run() { timeout ${TMO:-200} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline "$@" 2> gpurun_out/multi_err.log | tail -1 | python -c "
import json,sys
t=sys.stdin.read().strip()
try:
    d=json.loads(t); print('$N GPUs', ' '.join(sys.argv[1:]), '->', round(d['value']), 'samples/s', round(d['ms_per_step'],4), 'ms/step', 'e2e', round(d['e2e']['value']), '|', d['config'].get('allreduce')[:60])
    open('gpurun_out/scale_n${N}.jsonl','a').write(t+'\n')
except Exception as e: print('FAILED', e, t[:300])
" "$@"; grep -v "^\*\*\*\|OMP_NUM\|FutureWarning\|enable_symm" gpurun_out/multi_err.log | tail -3 | cut -c1-300; }
run --transport peer
VQA_PEER_THREADS=256 run --transport peer
VQA_PEER_THREADS=256 VQA_BUCKETS=6 run --transport peer
run --transport peer --batch 512
python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1 GPU batch 256 ->', round(d['value']), 'samples/s', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))"
