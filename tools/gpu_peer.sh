# N-GPU: data-parallel correctness with every transport, then bench lines: peer (NVLS / plain P2P) vs NCCL
N=${1:-2}
timeout 900 python -m pytest tests/test_dp_gpu.py -x -q -m gpu 2>&1 | tail -15
run() { timeout ${TMO:-150} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline "$@" 2> gpurun_out/multi_err.log | tail -1 | python -c "
import json,sys
t=sys.stdin.read().strip()
try:
    d=json.loads(t); print('$N GPUs', ' '.join(sys.argv[1:]), '->', round(d['value']), 'samples/s', round(d['ms_per_step'],4), 'ms/step', 'e2e', round(d['e2e']['value']), '|', d['config'].get('allreduce')[:110])
    open('gpurun_out/peer_n${N}.jsonl','a').write(t+'\n')
except Exception as e: print('FAILED', e, t[:300])
" "$@"; grep -v "^\*\*\*\|OMP_NUM\|FutureWarning\|enable_symm" gpurun_out/multi_err.log | tail -4 | cut -c1-300; }
run --transport peer
run --transport peer --no-overlap
VQA_PEER_MC=0 run --transport peer
run --transport nccl
