# end-of-round evidence: everything tools/gpu_evidence.sh collects + the inference sweep + the large ODA configuration
bash tools/gpu_evidence.sh ${1:-r1}
python tools/sweep_inference.py > gpurun_out/${1:-r1}_inference_sweep.md 2> gpurun_out/sweep.err; tail -26 gpurun_out/${1:-r1}_inference_sweep.md; tail -3 gpurun_out/sweep.err
for prec in tf32x3 tf32; do
python bench.py --model ODA --batch 512 --regions 100 --precision $prec --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${1:-r1}_bench_oda_b512_n100_$prec.json
python -c "
import json; d=json.loads(open('gpurun_out/${1:-r1}_bench_oda_b512_n100_$prec.json').read()); print('ODA 512x100 $prec', round(d['value']), 'samples/s', round(d['ms_per_step'],3), 'ms')"
done
python bench.py --batch 512 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/${1:-r1}_bench_cor2_b512.json
python -c "
import json; d=json.loads(open('gpurun_out/${1:-r1}_bench_cor2_b512.json').read()); print('CoR2 512x36', round(d['value']), 'samples/s', round(d['ms_per_step'],3), 'ms')"
