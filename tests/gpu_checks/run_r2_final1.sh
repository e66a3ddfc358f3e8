#!/bin/bash
# round 2, final 1-GPU call: launch list of a 4-layer Llama-3-8B step (ncu, duration only), then compute-sanitizer memcheck
mkdir -p gpurun_out
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2_4L.csv \
  python bench.py --layers 4 --global-batch 2 --steps 1 --warmup 1 --no-e2e > gpurun_out/final1_launch.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/launches_r2_4L.csv
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 7 python tests/gpu_checks/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1
echo "== memcheck exit $?"; grep -c "SANITIZE_SMALL_DONE" gpurun_out/sanitize_memcheck.log; grep "ERROR SUMMARY" gpurun_out/sanitize_memcheck.log | head -3
