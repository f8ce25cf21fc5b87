# Round evidence on the GPU box: (optional) parity suite, default bench (with CPU baseline), reference arm, ODA benches
# (256x36 fp32-parity; 512x100 in bf16 = BASELINE config 3, and in bf16x3), per-launch ncu list of one eager step of each
# model, `ncu --set full` of the dominant kernel (cor_compound_bwd), the attention-logits kernels and the compress GEMM.
#   bash tools/gpu_evidence.sh r2 [pytest]
TAG=${1:-r2}
if [ "$2" = "pytest" ]; then python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_gpu.txt; fi
python bench.py > gpurun_out/${TAG}_bench_cor2.json 2> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench_cor2.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench_reference.json; echo
python bench.py --model ODA --no-cpu-baseline > gpurun_out/${TAG}_bench_oda.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --model ODA --batch 512 --regions 100 --precision bf16 --steps 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_oda_b512_n100_bf16.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --model ODA --batch 512 --regions 100 --steps 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_oda_b512_n100_bf16x3.json 2>> gpurun_out/${TAG}_bench.err
for f in oda oda_b512_n100_bf16 oda_b512_n100_bf16x3; do python - <<EOF
import json
d=json.loads(open("gpurun_out/${TAG}_bench_$f.json").read().strip().splitlines()[-1])
print("$f", round(d["value"]), "samples/s", round(d["ms_per_step"],4), "ms/step e2e", round(d["e2e"]["value"]), d["roofline"]["kernel"][:40], round(d["roofline"]["frac"],3))
EOF
done
for m in CoR2 ODA; do
  ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/${TAG}_step_metrics_$m.csv \
      --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum python tools/ncu_step.py --model $m > gpurun_out/ncu_$m.log 2>&1
done
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"cor_compound_bwd|att_logits_softmax" -c 5 \
    -o gpurun_out/${TAG}_hbm_kernels -f python tools/ncu_step.py --model CoR2 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_hbm_kernels.ncu-rep --page details > gpurun_out/${TAG}_ncu_hbm_kernels_details.txt 2>/dev/null
ncu -i gpurun_out/${TAG}_hbm_kernels.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_ncu_hbm_kernels_source.csv 2>/dev/null
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:gemm16_kernel -c 1 \
    -o gpurun_out/${TAG}_gemm16 -f python tools/ncu_step.py --model CoR2 >> gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_gemm16.ncu-rep --page details > gpurun_out/${TAG}_ncu_gemm16_details.txt 2>/dev/null
tail -2 gpurun_out/ncu_full.log
