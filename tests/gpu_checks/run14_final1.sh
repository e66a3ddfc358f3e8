#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_native_32L_n1.log 2>&1; echo "exit $?" >> gpurun_out/bench_native_32L_n1.log; tail -n 2 gpurun_out/bench_native_32L_n1.log | cut -c1-1400
timeout 900 python bench.py --gpus 1 --steps 2 --warmup 3 --impl baseline > gpurun_out/bench_baseline_32L_n1.log 2>&1; echo "exit $?" >> gpurun_out/bench_baseline_32L_n1.log; tail -n 2 gpurun_out/bench_baseline_32L_n1.log | cut -c1-700
timeout 100 python bench.py --impl reference; 
