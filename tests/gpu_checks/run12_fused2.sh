#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 tests/gpu_checks/check_fused_tp.py > gpurun_out/check_fused_tp_n2.log 2>&1; echo "exit $?" >> gpurun_out/check_fused_tp_n2.log
grep -E "^AG|^GEMM|FUSED|^exit|Error|TIMELINE" gpurun_out/check_fused_tp_n2.log | head -30
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 tests/gpu_checks/check_fused_zero1.py > gpurun_out/check_fused_zero1_n2.log 2>&1; echo "exit $?" >> gpurun_out/check_fused_zero1_n2.log
grep -E "^step|ZERO1|^exit|Error|error" gpurun_out/check_fused_zero1_n2.log | head
