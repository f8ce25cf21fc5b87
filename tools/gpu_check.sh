# GPU-box check: parity suite, default bench, per-launch durations of one eager step (ncu, cheap metrics)
TAG=${1:-r1x}
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_cor2.json 2> gpurun_out/${TAG}.err; python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_cor2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['per_op_ms'])
PY
ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/${TAG}_step_metrics_cor2.csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum python tools/ncu_step.py --model CoR2 > gpurun_out/ncu1.log 2>&1; tail -1 gpurun_out/ncu1.log
