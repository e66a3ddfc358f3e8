#!/bin/bash
mkdir -p gpurun_out
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3"
run() { name=$1; shift; timeout 420 $B "$@" > gpurun_out/n8_$name.log 2>&1; echo "== $name exit $?"; grep -E '^\{"metric"' gpurun_out/n8_$name.log | cut -c1-2200; grep -E "Error|error|OutOfMemory" gpurun_out/n8_$name.log | head -3 | cut -c1-300; }
run default
run overlap --overlap --no-e2e
run tp4dp2 --tp 4 --no-e2e
run 70b --model llama3_70b --no-e2e
