#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
for mode in pull push; do
NXDT_AG_MODE=$mode timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$((RANDOM % 9)) tests/gpu_checks/check_fused_tp.py > gpurun_out/check_fused_tp_n2_$mode.log 2>&1; echo "exit $?" >> gpurun_out/check_fused_tp_n2_$mode.log
echo "== $mode"; grep -E "^AG|^GEMM|FUSED|^exit|Error|TIMELINE" gpurun_out/check_fused_tp_n2_$mode.log | head -30
done
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "flash or qkv or cross" > gpurun_out/pytest_new.log 2>&1; tail -3 gpurun_out/pytest_new.log
