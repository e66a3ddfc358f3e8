#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8_final.log 2>&1
echo "== N=8 exit $?"
grep -h '^{"metric"' gpurun_out/bench_n8_final.log | cut -c1-1500
grep -E "nxdt\]|Error" gpurun_out/bench_n8_final.log | head -5 | cut -c1-200
