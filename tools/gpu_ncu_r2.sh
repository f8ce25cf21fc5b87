#!/bin/bash
# ncu evidence for round 2 (run under gpurun, 1 GPU): per-launch list of one eager CoR2 step + full capture of the top kernels
set -x
mkdir -p gpurun_out
P=${1:-bf16x3}
ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r2_step_metrics_cor2_$P.csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    python tools/ncu_step.py --model CoR2 --precision $P > gpurun_out/r2_ncu_step_$P.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm16_kernel -c 12 \
    -o gpurun_out/r2_gemm16_$P -f python tools/ncu_step.py --model CoR2 --precision $P > gpurun_out/r2_ncu_full_$P.log 2>&1
ls -la gpurun_out/*.ncu-rep
