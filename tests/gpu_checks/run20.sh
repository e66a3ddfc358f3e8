#!/bin/bash
# quick single-GPU checks: remaining kernel tests + CUDA-graph capture of the micro-batch on a 4-layer slice
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "dropout or embedding or eager" > gpurun_out/pytest_gpu_r20.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu_r20.log
for g in 0 1; do
  NXDT_CUDA_GRAPH_DEBUG=1 NXDT_CUDA_GRAPH=$g timeout 200 python bench.py --steps 3 --warmup 3 --seq-len 1024 --layers 8 --recompute none --no-e2e > gpurun_out/bench_n1_8L_graph$g.log 2>&1
  echo "8L s1024 graph=$g exit $?"; grep -h '^{"metric"' gpurun_out/bench_n1_8L_graph$g.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('tokens/s', round(d['value']), 'ms/step', round(d['ms_per_step'], 1), 'loss', d['config'].get('final_loss'), 'graph', d['config'].get('cuda_graph'), 'launches', d['gpu_launches'])"
  grep -v '^{"metric"' gpurun_out/bench_n1_8L_graph$g.log | tail -25 | cut -c1-400
done
