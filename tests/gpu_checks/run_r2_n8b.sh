#!/bin/bash
mkdir -p gpurun_out
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3"
run() { name=$1; shift; timeout 420 "$@" > gpurun_out/n8b_$name.log 2>&1; echo "== $name exit $?"; grep -E '^\{"metric"' gpurun_out/n8b_$name.log | cut -c1-1700; grep -E "Error|error|OutOfMemory" gpurun_out/n8b_$name.log | head -3 | cut -c1-300; }
run default $B
run fp8 $B --fp8 --no-e2e
run keepgather env NXDT_KEEP_GATHERED=1 $B --no-e2e
run llama2_7b_tp2pp2dp2 $B --model llama2_7b --tp 2 --pp 2 --seq-len 4096 --global-batch 32 --no-e2e
