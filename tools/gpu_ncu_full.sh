# one `ncu --set full` capture of the compress_v2 wgrad + dgrad launches and the compress_v forward of one eager CoR2 step
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:tc_gemm_kernel -s 29 -c 2 \
    -o gpurun_out/r1i_wgrad_dgrad -f python tools/ncu_step.py --model CoR2 > gpurun_out/ncu_full1.log 2>&1
tail -3 gpurun_out/ncu_full1.log
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:tc_gemm_kernel -s 2 -c 1 \
    -o gpurun_out/r1i_compress_fwd -f python tools/ncu_step.py --model CoR2 > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/ncu_full2.log
ls -la gpurun_out/*.ncu-rep
