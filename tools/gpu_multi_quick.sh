# one overlapped N-GPU bench line under a short timeout (checks that the ranks exit cleanly)
N=${1:-2}
start=$(date +%s)
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/multi_err.log | tail -1 | cut -c1-400
echo "exit=$? elapsed=$(( $(date +%s) - start ))s"; tail -2 gpurun_out/multi_err.log | cut -c1-200
