#!/bin/bash
mkdir -p gpurun_out
NXDT_FLASH_VARIANT=0 timeout 100 python tests/gpu_checks/dbg_window.py 2>&1 | tail -14
timeout 150 python tests/gpu_checks/check_flash_bwd.py 2>&1 | grep -E "TFLOP|OK|rror|rel err" | tail -6 | cut -c1-200
