#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/pytest_gpu_r18.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu_r18.log
timeout 300 python tests/gpu_checks/check_grouped.py > gpurun_out/check_grouped.log 2>&1; echo "grouped exit $?"; tail -8 gpurun_out/check_grouped.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1_default.log 2>&1; echo "bench exit $?"; grep -h '^{"metric"' gpurun_out/bench_n1_default.log | cut -c1-900; tail -3 gpurun_out/bench_n1_default.log | cut -c1-300
