#!/bin/bash
# ncu captures (ONE GPU; ncu replays every kernel ~40x, so the targets launch each kernel twice) for the kernels without a
# committed profile.  Bring the .ncu-rep files back in gpurun_out/ and summarise them here with tools/ncu_summary.py:
#   gpurun --timeout 900 -- 'bash tests/gpu_checks/run_round2_ncu.sh flash_fwd flash_fwd_split2 grouped'
#   python tools/ncu_summary.py gpurun_out/prof_flash_fwd.ncu-rep > profiles/flash_fwd_sm100.md
mkdir -p gpurun_out
for t in "${@:-flash_fwd}"; do
  envs=""; target=$t; kern="regex:flash_fwd"
  case $t in
    flash_fwd) ;;
    flash_fwd_split2) envs="NXDT_FLASH_SPLIT=2"; target=flash_fwd ;;
    flash_bwd) kern="regex:flash_bwd_kernel" ;;
    fp8) kern="regex:gemm_fp8_kernel" ;;
    grouped) kern="regex:grouped_rows_kernel" ;;
    *) echo "unknown target $t"; continue ;;
  esac
  env $envs timeout 600 ncu --set full --clock-control none --import-source on -k "$kern" -c 1 -s 1 -f \
      -o gpurun_out/prof_$t python tests/gpu_checks/prof_kernels.py $target > gpurun_out/ncu_$t.log 2>&1
  echo "== $t exit $? $(ls -la gpurun_out/prof_$t.ncu-rep 2>/dev/null | awk '{print $5}') bytes; $(grep -h 'PROF_TARGET_DONE\|Error\|error' gpurun_out/ncu_$t.log | tail -2 | tr '\n' '|')"
done
