#!/bin/bash
mkdir -p gpurun_out
for v in 0 t 1 2; do
  echo "=== variant $v"
  NXDT_FLASH_VARIANT=$v timeout 100 python tests/gpu_checks/check_flash.py 2>&1 | grep -E "8B shape|non-causal B1 S2048|w700|Error|error" | cut -c1-200
  NXDT_FLASH_VARIANT=$v timeout 60 python tests/gpu_checks/timeline_flash.py 2>&1 | tail -1
done
