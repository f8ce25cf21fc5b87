#!/bin/bash
# compute-sanitizer over one small train step (batch 32): memcheck + racecheck of the ODA plan (the rebuilt pairwise
# kernels: shared tile, cp.async ring, named barriers) and memcheck of the CoR2 plan.  Bounded by timeouts.
mkdir -p gpurun_out
run() { tool=$1; model=$2; out=gpurun_out/sanitize_${tool}_${model}.txt
  timeout ${TMO:-200} compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 10 \
    python tools/ncu_step.py --model $model --batch 32 --steps 1 > $out 2>&1
  echo "$tool $model exit=$?" | tee -a gpurun_out/sanitize_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|error" $out | tail -4 | cut -c1-200 | tee -a gpurun_out/sanitize_summary.txt; }
[ "$KEEP" = 1 ] || rm -f gpurun_out/sanitize_summary.txt
run racecheck ODA
run memcheck ODA
run memcheck CoR2
run racecheck CoR2
run synccheck CoR2
run synccheck ODA
