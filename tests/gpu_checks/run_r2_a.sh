#!/bin/bash
# round 2, call A (1 GPU): validate kernels that were compiled but never executed in round 1.
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; timeout "$t" "$@" > gpurun_out/r2_$name.log 2>&1
  echo "== $name exit $? : $(grep -E '_OK|rel err|rel_err|TFLOP|TB/s|passed|failed|Error|error|assert' gpurun_out/r2_$name.log | tail -6 | cut -c1-200 | tr '\n' '|')"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
run flash_bwd 150 python tests/gpu_checks/check_flash_bwd.py
run flash_fwd 120 python tests/gpu_checks/check_flash.py
NXDT_FLASH_SPLIT=2 run flash_fwd_split2 120 python tests/gpu_checks/check_flash.py
run fp8 180 python tests/gpu_checks/check_fp8.py
run adam_v1 120 python tests/gpu_checks/check_adam.py
NXDT_ADAM_V2=1 run adam_v2 120 python tests/gpu_checks/check_adam.py
run smoke 120 python -c "import __graft_entry__ as g; g.smoke()"
