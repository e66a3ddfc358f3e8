#!/bin/bash
# round 2, call C (1 GPU): SM-pair GEMM as the default route -> layout probe, N=1 bench, GPU test suite, ncu captures
mkdir -p gpurun_out
GEMM_IMPL=2cta timeout 240 python tests/gpu_checks/check_gemm.py 2>&1 | grep "bench\|acc" > gpurun_out/c_gemm2cta.log; tail -n 12 gpurun_out/c_gemm2cta.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/c_bench_n1.log 2>&1; echo "bench rc=$?"; grep '^{"metric' gpurun_out/c_bench_n1.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/c_pytest.log; cat gpurun_out/c_pytest.log
bash tests/gpu_checks/run_r2_ncu.sh
