#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/gpu_checks/check_fused_tp.py > gpurun_out/check_fused_tp_n8.log 2>&1; echo "exit $?" >> gpurun_out/check_fused_tp_n8.log
grep -E "^AG|^GEMM|FUSED|^exit|TIMELINE" gpurun_out/check_fused_tp_n8.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 2 --warmup 2 > gpurun_out/bench_fused_32L_n8.log 2>&1; echo "exit $?" >> gpurun_out/bench_fused_32L_n8.log
tail -n 3 gpurun_out/bench_fused_32L_n8.log | cut -c1-600
