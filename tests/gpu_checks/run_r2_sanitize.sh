#!/bin/bash
# compute-sanitizer over the small-shape launches of the tcgen05 / TMA kernels (1 GPU)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 7 python tests/gpu_checks/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool exit $?"; grep -c "SANITIZE_SMALL_DONE" gpurun_out/sanitize_$tool.log; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|=========.*[Ee]rror\|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -8
done
