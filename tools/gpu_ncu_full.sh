# `ncu --set full` captures of selected tc_gemm launches of one eager CoR2 step: $1 = first launch index, $2 = count, $3 = tag
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:tc_gemm_kernel -s $1 -c $2 \
    -o gpurun_out/$3 -f python tools/ncu_step.py --model CoR2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
