#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; timeout "$t" "$@" > gpurun_out/r2_$name.log 2>&1
  echo "== $name exit $? : $(grep -E '_OK|rel err|rel_err|TFLOP|TB/s|passed|failed|Error|error|assert|timed out' gpurun_out/r2_$name.log | tail -${TAILN:-4} | cut -c1-220 | tr '\n' '|')"; }
run flash_fwd2 150 python tests/gpu_checks/check_flash.py
run flash_bwd2 150 python tests/gpu_checks/check_flash_bwd.py
