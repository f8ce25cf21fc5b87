N=${1:-2}
VQA_ALLREDUCE=${2:-peer} timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 tools/dp_debug.py eval_cap train_cap train_cap_opt train_nocap_opt 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|FutureWarning\|enable_symm" | tail -50
