#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/gpu_checks/check_flash.py > gpurun_out/check_flash.log 2>&1; echo "exit $?" >> gpurun_out/check_flash.log
cat gpurun_out/check_flash.log | tail -12
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_kernels.log
tail -4 gpurun_out/pytest_kernels.log
bash tests/gpu_checks/run8_prof.sh
