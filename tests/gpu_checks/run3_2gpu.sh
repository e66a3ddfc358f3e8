#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/bench_native_32L_n2.log 2>&1; echo "exit $?" >> gpurun_out/bench_native_32L_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 2 --warmup 2 --impl baseline > gpurun_out/bench_baseline_32L_n2.log 2>&1; echo "exit $?" >> gpurun_out/bench_baseline_32L_n2.log
for f in bench_native_32L_n2 bench_baseline_32L_n2; do echo "== $f"; tail -n 5 gpurun_out/$f.log | cut -c1-1500; done
