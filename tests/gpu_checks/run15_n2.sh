#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 tests/gpu_checks/check_fused_zero1.py > gpurun_out/check_fused_zero1_n2.log 2>&1; echo "exit $?" >> gpurun_out/check_fused_zero1_n2.log
grep -E "^step|ZERO1|^exit|Error" gpurun_out/check_fused_zero1_n2.log | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_native_32L_n2.log 2>&1; echo "exit $?" >> gpurun_out/bench_native_32L_n2.log
tail -n 2 gpurun_out/bench_native_32L_n2.log | cut -c1-600
