set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r1h_bench_cor2.json 2> gpurun_out/r1h_bench_cor2.err; tail -c 3000 gpurun_out/r1h_bench_cor2.json
ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r1h_step_metrics_cor2.csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum python tools/ncu_step.py --model CoR2 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r1h_step_metrics_oda.csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum python tools/ncu_step.py --model ODA > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
