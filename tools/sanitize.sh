#!/usr/bin/env bash
# Race / memory checking of the sm_100a kernels.  Run on a GPU box:
#   tools/sanitize.sh memcheck|racecheck|synccheck|initcheck [pytest -k expression]
# Cross-GPU flag / barrier protocols of the fused kernels are exercised by tests/gpu_checks/check_fused_tp.py
# (run it under `compute-sanitizer --tool memcheck` with torchrun --nproc-per-node 2 for peer-access errors).
set -euo pipefail
TOOL=${1:-memcheck}
EXPR=${2:-"gemm or rmsnorm or rope or swiglu or adamw or cross_entropy or qkv or flash"}
cd "$(dirname "$0")/.."
exec compute-sanitizer --tool "$TOOL" --error-exitcode 1 --target-processes all \
  python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "$EXPR"
