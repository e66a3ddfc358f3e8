#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 2 --warmup 2 > gpurun_out/bench_native_32L_n$n.log 2>&1; echo "exit $?" >> gpurun_out/bench_native_32L_n$n.log
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 2 --warmup 2 --impl baseline > gpurun_out/bench_baseline_32L_n8.log 2>&1; echo "exit $?" >> gpurun_out/bench_baseline_32L_n8.log
for f in bench_native_32L_n8 bench_native_32L_n4 bench_baseline_32L_n8; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-1200; done
