"""Build / load the native checkpoint IO extension (``checkpoint/csrc/ckpt_io.cpp``)."""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc", "ckpt_io.cpp")
_SO = os.path.join(_HERE, "_ckpt_io.so")
_MOD = None
_TRIED = False


def build(force: bool = False) -> str:
    import shutil
    from torch.utils import cpp_extension
    deps = [_SRC, os.path.join(os.path.dirname(_SRC), "parallel_io.h")]
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= max(os.path.getmtime(d) for d in deps):
        return _SO
    bdir = os.path.join(_HERE, "_build")
    os.makedirs(bdir, exist_ok=True)
    cpp_extension.load(name="_ckpt_io", sources=[_SRC], build_directory=bdir, extra_cflags=["-O3", "-std=c++17"],
                       is_python_module=True, with_cuda=False, verbose=False)
    shutil.copy2(os.path.join(bdir, "_ckpt_io.so"), _SO)
    return _SO


class _Wrapper:
    def __init__(self, mod):
        self.mod = mod

    def write_file(self, path, header: bytes, raw_u8):
        self.mod.write_file(path, header, raw_u8, 4, False)

    def read_file(self, path, offset: int, out_u8):
        self.mod.read_file(path, int(offset), out_u8, 4)


def get():
    global _MOD, _TRIED
    if _MOD is not None or _TRIED:
        return _MOD
    _TRIED = True
    if not os.path.exists(_SO):
        return None
    try:
        import torch  # noqa: F401
        loader = importlib.machinery.ExtensionFileLoader("_ckpt_io", _SO)
        spec = importlib.util.spec_from_file_location("_ckpt_io", _SO, loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _MOD = _Wrapper(mod)
    except Exception:
        _MOD = None
    return _MOD
