"""Loader / builder for the in-tree sm_100a extension ``ops/_C*.so``.

The extension is compiled ahead of time (``build()``; nvcc cross-compiles sm_100a without a GPU) and
lives in-tree so it travels to the GPU box with the repo snapshot.  ``get()`` only ever *loads* it.
"""
from __future__ import annotations

import glob
import importlib.machinery
import importlib.util
import os
from typing import List, Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_MOD = None
_ERR: Optional[BaseException] = None
_TRIED = False

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "--expt-extended-lambda", "-Xptxas", "-v",
    "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
    "-U__CUDA_NO_BFLOAT16_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__",
]


def sources() -> List[str]:
    return sorted(glob.glob(os.path.join(_CSRC, "*.cu")) + glob.glob(os.path.join(_CSRC, "*.cpp")))


def so_path() -> Optional[str]:
    c = sorted(glob.glob(os.path.join(_HERE, "_C*.so")))
    return c[0] if c else None


def build(verbose: bool = False, force: bool = False) -> str:
    """Compile every .cu/.cpp under ops/csrc into ops/_C.so (ninja via torch cpp_extension)."""
    import torch  # noqa
    from torch.utils import cpp_extension
    build_dir = os.path.join(_HERE, "_build")
    os.makedirs(build_dir, exist_ok=True)
    srcs = sources()
    target = os.path.join(_HERE, "_C.so")
    if not force and os.path.exists(target):
        newest = max(os.path.getmtime(p) for p in srcs + glob.glob(os.path.join(_CSRC, "*.cuh")) +
                     glob.glob(os.path.join(_CSRC, "*.h")))
        if os.path.getmtime(target) >= newest:
            return target
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    cpp_extension.load(
        name="_C", sources=srcs, build_directory=build_dir, verbose=verbose,
        extra_cflags=["-O3", "-std=c++17"], extra_cuda_cflags=NVCC_FLAGS,
        extra_ldflags=["-lcuda"], is_python_module=True, with_cuda=True,
    )
    built = os.path.join(build_dir, "_C.so")
    if not os.path.exists(built):
        raise RuntimeError("extension build produced no _C.so")
    import shutil
    shutil.copy2(built, target)
    global _MOD, _TRIED, _ERR
    _MOD, _TRIED, _ERR = None, False, None
    return target


def get():
    """The loaded extension module, or None (with ``load_error()`` set)."""
    global _MOD, _ERR, _TRIED
    if _MOD is not None or _TRIED:
        return _MOD
    _TRIED = True
    path = so_path()
    if path is None:
        _ERR = FileNotFoundError(f"no _C*.so under {_HERE}")
        return None
    try:
        import torch  # noqa: F401  (libtorch symbols must be loaded first)
        loader = importlib.machinery.ExtensionFileLoader("_C", path)
        spec = importlib.util.spec_from_file_location("_C", path, loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _MOD = mod
    except BaseException as e:  # noqa
        _ERR = e
        _MOD = None
    return _MOD


def load_error():
    return _ERR


def is_available() -> bool:
    return get() is not None
