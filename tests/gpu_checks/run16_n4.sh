#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
FUSED_TIMELINE=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 tests/gpu_checks/check_fused_tp.py > gpurun_out/check_fused_tp_n4.log 2>&1; echo "exit $?" >> gpurun_out/check_fused_tp_n4.log
grep -E "^AG|^GEMM|FUSED|^exit|Error" gpurun_out/check_fused_tp_n4.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_native_32L_n4.log 2>&1; echo "exit $?" >> gpurun_out/bench_native_32L_n4.log
tail -n 2 gpurun_out/bench_native_32L_n4.log | cut -c1-600
