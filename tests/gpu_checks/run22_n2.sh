#!/bin/bash
# N=2: graph capture with fused TP after serialising NCCL handles before fused kernels (4-layer slice), then the
# default (eager) full-model bench to confirm no regression.
mkdir -p gpurun_out
export NCCL_DEBUG=WARN PYTHONFAULTHANDLER=1 NXDT_CUDA_GRAPH_DEBUG=1
N=2
NXDT_CUDA_GRAPH=1 timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 2 --warmup 3 --layers 4 --no-e2e > gpurun_out/diag2_graph_fused.log 2>&1
echo "== graph_fused 4L exit $?"
grep -h '^{"metric"' gpurun_out/diag2_graph_fused.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('tokens/s', round(d['value']), 'ms/step', round(d['ms_per_step'], 1), 'loss', d['config'].get('final_loss'), 'graph', d['config'].get('cuda_graph'), 'launches', d.get('gpu_launches'))"
grep -E "nxdt\]|Fatal Python|CUDA-graph|Error" gpurun_out/diag2_graph_fused.log | head -8 | cut -c1-250
timeout 160 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n2_final.log 2>&1
echo "== eager 32L exit $?"
grep -h '^{"metric"' gpurun_out/bench_n2_final.log | cut -c1-1200
