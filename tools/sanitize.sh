#!/bin/bash
# compute-sanitizer over one train step of both models (run under gpurun, 1 GPU):
#   memcheck  - out-of-bounds / misaligned accesses of every kernel, the TMA-fed tcgen05 GEMMs included
#   racecheck - shared-memory hazards (the hand-rolled mbarrier pipelines and the epilogue slabs live there)
#   synccheck - divergent barriers (named barriers of the epilogue warps)
# A small batch keeps the instrumented run short; B = 32 x 36 regions still reaches the large-M bf16-plane kernel.
# Summaries land in gpurun_out/sanitize_<tool>_<model>.txt; profiles/ keeps the last lines.
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for model in CoR2 ODA; do
    out=gpurun_out/sanitize_${tool}_${model}.txt
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 \
      python tools/ncu_step.py --model $model --batch 32 --precision ${PRECISION:-bf16x3} > $out 2>&1
    echo "$tool $model exit=$?" | tee -a gpurun_out/sanitize_summary.txt
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $out | tail -3 | tee -a gpurun_out/sanitize_summary.txt
  done
done
