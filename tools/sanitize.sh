#!/bin/bash
# compute-sanitizer over the committed kernels at small sizes (memcheck / racecheck on the atomics
# paths).  Run under gpurun; summaries land in gpurun_out/sanitizer_*.txt (copy what is kept into profiles/).
set -u
export CB200_PROFILE_SMALL=1
mkdir -p gpurun_out
for what in loss sampled detect post; do
  for tool in memcheck racecheck; do
    out=gpurun_out/sanitizer_${tool}_${what}.txt
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/profile_target.py $what > $out 2>&1
    echo "== $tool $what: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out | tail -1)"
  done
done
