#!/bin/bash
# round 2, call F (2 GPUs): NVLS kernels (tuned), the whole-model NVLS parity tests, the previously flaky whole-model test in suite order
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 \
  tests/gpu_checks/check_nvls.py > gpurun_out/f_nvls.log 2>&1; echo "nvls rc=$?"; grep -v "^W\|^\[W\|warn\|^\*\|OMP_NUM" gpurun_out/f_nvls.log | tail -n 16
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -k "nvls" > gpurun_out/f_pytest_nvls.log 2>&1; echo "pytest nvls rc=$?"; tail -n 30 gpurun_out/f_pytest_nvls.log | cut -c1-300
timeout 900 python -m pytest tests/test_families_gpu.py tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/f_pytest_suite.log 2>&1; echo "pytest suite rc=$?"; tail -n 5 gpurun_out/f_pytest_suite.log | cut -c1-300
