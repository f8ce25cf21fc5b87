# Round evidence on the GPU box: parity suite, default bench (with CPU baseline), reference arm, ODA bench,
# per-launch ncu list of one eager step of each model, one `ncu --set full` capture of the dominant kernel.
TAG=${1:-r1}
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_gpu.txt
python bench.py > gpurun_out/${TAG}_bench_cor2.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench_cor2.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench_reference.json
python bench.py --model ODA --no-cpu-baseline > gpurun_out/${TAG}_bench_oda.json 2>> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench_oda.json
for m in CoR2 ODA; do
  ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/${TAG}_step_metrics_$m.csv \
      --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum python tools/ncu_step.py --model $m > gpurun_out/ncu_$m.log 2>&1
done
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:tc_gemm_kernel -s 29 -c 2 \
    -o gpurun_out/${TAG}_wgrad_dgrad -f python tools/ncu_step.py --model CoR2 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_wgrad_dgrad.ncu-rep --page details > gpurun_out/${TAG}_ncu_wgrad_dgrad_details.txt 2>/dev/null
tail -2 gpurun_out/ncu_full.log
