#!/bin/bash
mkdir -p gpurun_out
timeout 60 python tests/gpu_checks/timeline_flash_bwd.py 2>&1 | tail -22 | cut -c1-260
timeout 1200 python -m pytest tests/test_multigpu.py -x -q -m gpu -s 2>&1 | grep -vE "^\s*$" | tail -40 | cut -c1-330
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e"
timeout 400 $B 2>&1 | grep -E '^\{"metric"|Error|error' | cut -c1-1500
timeout 400 $B --overlap 2>&1 | grep -E '^\{"metric"|Error|error' | cut -c1-1500
