python -m pytest tests/test_model_parity.py -x -q -m gpu -k "ODA or oda" 2>&1 | tail -3
python bench.py --model ODA --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ODA', round(d['value']), 'samples/s', round(d['ms_per_step'],4), 'ms', list(d['per_op_ms'].items())[:4])"
ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/oda_step.csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum python tools/ncu_step.py --model ODA > /dev/null 2>&1
python tools/ncu_table.py gpurun_out/oda_step.csv | grep -E "total|oda_"
