#!/bin/bash
# Round-2 ncu captures (one GPU): every headline kernel once, --set full, clocks untouched.  Reports come back in gpurun_out/.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'flash_fwd2_kernel|flash_bwd2_kernel' \
  -f -o gpurun_out/r2_flash python tests/gpu_checks/prof_kernels.py > gpurun_out/r2_ncu_flash.log 2>&1
echo "flash rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^gemm_kernel|gemm2cta_kernel|gemm_mxfp8_kernel|mx_quantize' \
  -f -o gpurun_out/r2_gemm python tests/gpu_checks/prof_kernels.py > gpurun_out/r2_ncu_gemm.log 2>&1
echo "gemm rc=$?"
ls -la gpurun_out/*.ncu-rep
