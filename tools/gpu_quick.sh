python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for m in ODA CoR2; do python bench.py --model $m --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m', round(d['value']), 'samples/s', round(d['ms_per_step'],4), 'ms', list(d['per_op_ms'].items())[:6])"; done
