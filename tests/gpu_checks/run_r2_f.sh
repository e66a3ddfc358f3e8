#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tests/gpu_checks/timeline_flash.py 2>&1 | tail -24
