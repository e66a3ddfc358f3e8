#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 | cut -c1-300
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e --layers 8 --recompute none"
timeout 400 $B --tp 1 --pp 2 2>&1 | grep -E '^\{"metric"|Error|error' | cut -c1-1300
timeout 400 $B --tp 2 2>&1 | grep -E '^\{"metric"|Error|error' | cut -c1-1300
