#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tests/gpu_checks/check_flash.py 2>&1 | grep -E "8B shape|rror|rel_err" | cut -c1-170 | tail -16
timeout 200 python tests/gpu_checks/check_flash_bwd.py 2>&1 | grep -E "TFLOP|OK|rror|rel err" | tail -14 | cut -c1-200
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | cut -c1-300
