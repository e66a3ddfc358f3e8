#!/usr/bin/env bash
# One-time setup on a B200 machine (the reference's install_setup.sh builds the Megatron dataset helpers and patches
# NeMo; here everything native lives in this repository and is compiled in-tree for sm_100a).
set -euo pipefail
cd "$(dirname "$0")"
python -m pip install -e . --no-build-isolation --no-deps
python setup.py build_native          # ops/_C.so (nvcc -gencode arch=compute_100a,code=sm_100a), data helpers, checkpoint IO
python - <<'PY'
from neuronx_distributed_training_b200 import ops
print("sm_100a extension:", "ok" if ops.C() is not None else f"missing ({ops._native_loader.load_error()})")
PY
