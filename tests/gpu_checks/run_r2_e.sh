#!/bin/bash
# round 2, call E (2 GPUs): NVLS probe + kernels, then the whole GPU suite (single- and multi-GPU tests) with the SM-pair GEMM default
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
  tests/gpu_checks/check_nvls.py > gpurun_out/e_nvls.log 2>&1; echo "nvls rc=$?"; grep -v "^W\|^\[W\|warn" gpurun_out/e_nvls.log | tail -n 25
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 60 gpurun_out/e_pytest.log | cut -c1-400
