#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flash_fwd2_kernel -s 1 -c 1 -o gpurun_out/prof_flash_fwd2 -f python tests/gpu_checks/prof_flash.py > gpurun_out/ncu_fwd2.log 2>&1; echo "fwd2 ncu exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flash_bwd2_kernel -s 1 -c 1 -o gpurun_out/prof_flash_bwd2 -f python tests/gpu_checks/prof_flash.py > gpurun_out/ncu_bwd2.log 2>&1; echo "bwd2 ncu exit $?"
ls -la gpurun_out/*.ncu-rep
