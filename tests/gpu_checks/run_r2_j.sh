#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tests/gpu_checks/dbg_window.py 2>&1 | tail -12 | cut -c1-200
timeout 150 python tests/gpu_checks/check_flash.py 2>&1 | grep -E "8B shape|rror|rel_err" | awk '{print $0}' | cut -c1-160 | tail -14
timeout 150 python tests/gpu_checks/check_flash_bwd.py 2>&1 | grep -E "TFLOP|OK|rror|rel err" | tail -12 | cut -c1-200
timeout 60 python tests/gpu_checks/timeline_flash.py 2>&1 | tail -1
