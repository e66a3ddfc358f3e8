#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --impl native --steps 2 --warmup 2 > gpurun_out/bench_native_32L.log 2>&1; echo "exit $?" >> gpurun_out/bench_native_32L.log
timeout 900 python bench.py --impl baseline --steps 2 --warmup 2 > gpurun_out/bench_baseline_32L.log 2>&1; echo "exit $?" >> gpurun_out/bench_baseline_32L.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 3 -o gpurun_out/prof_gemm -f python tests/gpu_checks/prof_gemm.py > gpurun_out/ncu_gemm.log 2>&1; echo "exit $?" >> gpurun_out/ncu_gemm.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv --log-file gpurun_out/launches_native_4L.csv python bench.py --impl native --layers 4 --steps 1 --warmup 3 --global-batch 1 --no-e2e > gpurun_out/ncu_launches.log 2>&1; echo "exit $?" >> gpurun_out/ncu_launches.log
for f in bench_native_32L bench_baseline_32L ncu_gemm ncu_launches; do echo "== $f"; tail -n 4 gpurun_out/$f.log; done
