#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/gpu_checks/check_fused_tp.py > gpurun_out/check_fused_tp_n2.log 2>&1; echo "exit $?" >> gpurun_out/check_fused_tp_n2.log
grep -E "^AG|^GEMM|FUSED|^exit|Error|TIMELINE" gpurun_out/check_fused_tp_n2.log | head -30
