#!/bin/bash
# N=2 diagnosis of CUDA-graph capture with NCCL-only TP, fused TP, and without expandable segments (4-layer slice)
mkdir -p gpurun_out
export NCCL_DEBUG=WARN PYTHONFAULTHANDLER=1 NXDT_CUDA_GRAPH_DEBUG=1 NXDT_CUDA_GRAPH=1
N=2
run() {  # name, extra env...
  name=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 2 --warmup 3 --layers 4 --no-e2e > gpurun_out/diag_$name.log 2>&1
  echo "== $name exit $?"
  grep -h '^{"metric"' gpurun_out/diag_$name.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('tokens/s', round(d['value']), 'ms/step', round(d['ms_per_step'], 1), 'loss', d['config'].get('final_loss'), 'graph', d['config'].get('cuda_graph'), 'launches', d.get('gpu_launches'))"
  grep -E "nxdt\]|Fatal Python|File \"/tmp|File \"/root|Segmentation|CUDA-graph|Error" gpurun_out/diag_$name.log | head -30 | cut -c1-250
}
run eager_fused NXDT_CUDA_GRAPH=0
run graph_nccl NXDT_FUSED_TP=0
run graph_fused NXDT_FUSED_TP=1
run graph_fused_noexp NXDT_FUSED_TP=1 PYTORCH_CUDA_ALLOC_CONF=expandable_segments:False
