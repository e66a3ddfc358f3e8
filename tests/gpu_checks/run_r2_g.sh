#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; timeout "$t" "$@" > gpurun_out/r2_$name.log 2>&1
  echo "== $name exit $? : $(grep -E '_OK|rel err|rel_err|TFLOP|TB/s|passed|failed|Error|error|assert|timed out' gpurun_out/r2_$name.log | tail -${TAILN:-3} | cut -c1-220 | tr '\n' '|')"; }
TAILN=16 run flash_fwd2 150 python tests/gpu_checks/check_flash.py
timeout 120 python tests/gpu_checks/timeline_flash.py 2>&1 | tail -24
