#!/bin/bash
# 8-GPU comparison of the TP=8 step-time levers (each bench run costs ~10 GPU-minutes of budget: 8 GPUs x ~75 s).
#   gpurun --gpus 8 --timeout 900 -- 'bash tests/gpu_checks/run_round2_n8.sh default keepgather'
# modes: default | keepgather | overlap | graphnccl | nofuse | keepgather+overlap | tp4dp2 | tp4dp2_nccl_zero1
mkdir -p gpurun_out
export NCCL_DEBUG=WARN PYTHONFAULTHANDLER=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
port=29800
for mode in "${@:-default}"; do
  port=$((port + 1))
  envs=""; extra=""
  case $mode in
    default) ;;
    keepgather) envs="NXDT_KEEP_GATHERED=1" ;;
    overlap) envs="NXDT_MB_OVERLAP=1" ;;
    keepgather+overlap) envs="NXDT_KEEP_GATHERED=1 NXDT_MB_OVERLAP=1" ;;
    graphnccl) envs="NXDT_CUDA_GRAPH=1 NXDT_FUSED_TP=0" ;;
    nofuse) envs="NXDT_FUSED_TP=0" ;;
    tp4dp2) envs="" ; extra="--tp 4" ;;          # same model and batch, TP 4 x DP 2 (fused ZeRO-1 step): NOT the named config, a data point
    tp4dp2_nccl_zero1) envs="NXDT_FUSED_ZERO1=0" ; extra="--tp 4" ;;
    *) echo "unknown mode $mode"; continue ;;
  esac
  env $envs timeout 240 $TR --master-port $port bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e $extra > gpurun_out/r2n8_bench_$mode.log 2>&1
  echo "== bench $mode exit $?"
  grep -h '^{"metric"' gpurun_out/r2n8_bench_$mode.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); c = d['config']
    print('tokens/s', round(d['value']), 'ms/step', round(d['ms_per_step'], 1), 'loss', c.get('final_loss'), 'mem', c.get('peak_mem_gb'),
          'fused', c.get('fused_tp_layers'), 'keep', c.get('keep_gathered_input'), 'overlap', c.get('microbatch_overlap'), 'graph', c.get('cuda_graph'), 'par', c.get('parallelism'), 'fz1', c.get('fused_zero1'))"
  grep -h "nxdt\]\|Error\|error:" gpurun_out/r2n8_bench_$mode.log | head -3
done
