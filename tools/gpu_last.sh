python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/last_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/last_bench_cor2.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/last_bench_cor2.json').read().strip().splitlines()[-1]); print('CoR2', round(d['value']), 'samples/s', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value'],1), 'roof', round(d['roofline']['frac'],4), d['gpu_launches'])"
python bench.py --model ODA --no-cpu-baseline > gpurun_out/last_bench_oda.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/last_bench_oda.json').read().strip().splitlines()[-1]); print('ODA', round(d['value']), 'samples/s', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))"
