#!/bin/bash
# first GPU round: GEMM probe, kernel tests, smoke, short benches.  Every step has its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 240 python tests/gpu_checks/check_gemm.py > gpurun_out/check_gemm.log 2>&1; echo "check_gemm exit $?" >> gpurun_out/check_gemm.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_kernels.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --impl baseline --layers 4 --steps 2 --warmup 2 --global-batch 2 > gpurun_out/bench_baseline_4L.log 2>&1; echo "exit $?" >> gpurun_out/bench_baseline_4L.log
timeout 600 python bench.py --impl native --layers 4 --steps 2 --warmup 2 --global-batch 2 > gpurun_out/bench_native_4L.log 2>&1; echo "exit $?" >> gpurun_out/bench_native_4L.log
tail -5 gpurun_out/check_gemm.log gpurun_out/pytest_kernels.log gpurun_out/smoke.log gpurun_out/bench_baseline_4L.log gpurun_out/bench_native_4L.log
