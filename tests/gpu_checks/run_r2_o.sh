#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tests/gpu_checks/check_flash_bwd.py 2>&1 | grep -E "TFLOP|OK|rror|rel err" | tail -12 | cut -c1-200
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | cut -c1-300
