# A/B of an environment switch on the default bench: bash tools/gpu_ab.sh VAR valA valB
for v in $2 $3; do
  env $1=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1=$v', round(d['value']), 'samples/s', round(d['ms_per_step'],4), 'ms', {k:v for k,v in list(d['per_op_ms'].items())[:8]})
"
done
