#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/gpu_checks/check_mxfp8.py 2>&1 | tail -12 | cut -c1-260
