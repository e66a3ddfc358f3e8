#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/gpu_checks/check_fused_tp.py > gpurun_out/check_fused_tp_n2.log 2>&1; echo "exit $?" >> gpurun_out/check_fused_tp_n2.log
grep -E "^AG|^GEMM|FUSED|^exit|Error" gpurun_out/check_fused_tp_n2.log | head -20
NXDT_FUSE_THRESHOLD=0.0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 2 --warmup 2 --layers 8 --no-e2e > gpurun_out/bench_fusedall_8L_n2.log 2>&1; echo "exit $?" >> gpurun_out/bench_fusedall_8L_n2.log
tail -n 2 gpurun_out/bench_fusedall_8L_n2.log | cut -c1-400
NXDT_FUSE_THRESHOLD=9.0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 2 --layers 8 --no-e2e > gpurun_out/bench_nofuse_8L_n2.log 2>&1; echo "exit $?" >> gpurun_out/bench_nofuse_8L_n2.log
tail -n 2 gpurun_out/bench_nofuse_8L_n2.log | cut -c1-400
