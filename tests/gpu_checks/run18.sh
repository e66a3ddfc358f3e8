#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/pytest_gpu_r18.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu_r18.log
timeout 300 python tests/gpu_checks/check_grouped.py > gpurun_out/check_grouped.log 2>&1; echo "grouped exit $?"; tail -8 gpurun_out/check_grouped.log
# launch-bound probe: short sequences, eager vs CUDA-graph replay of the micro-batch (same losses expected)
for g in 0 1; do
  NXDT_CUDA_GRAPH=$g timeout 300 python bench.py --steps 3 --warmup 3 --seq-len 512 --recompute none --no-e2e > gpurun_out/bench_n1_s512_graph$g.log 2>&1
  echo "s512 graph=$g exit $?"; grep -h '^{"metric"' gpurun_out/bench_n1_s512_graph$g.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('tokens/s', round(d['value']), 'ms/step', round(d['ms_per_step'], 1), 'loss', d['config'].get('final_loss'), 'mem', d['config'].get('peak_mem_gb'))"
  tail -2 gpurun_out/bench_n1_s512_graph$g.log | cut -c1-300
done
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1_default.log 2>&1; echo "bench exit $?"; grep -h '^{"metric"' gpurun_out/bench_n1_default.log | cut -c1-900; tail -3 gpurun_out/bench_n1_default.log | cut -c1-300
