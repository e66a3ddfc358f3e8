"""Build / load the C++ index-builder extension (``data/csrc/helpers.cpp``) in-tree with g++ + pybind11."""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import subprocess
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc", "helpers.cpp")
_MOD = None


def so_path() -> str:
    return os.path.join(_HERE, "data_helpers" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build(force: bool = False) -> str:
    import pybind11
    out = so_path()
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(_SRC):
        return out
    inc = [pybind11.get_include(), sysconfig.get_paths()["include"]]
    cmd = ["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-Wall"] + [f"-I{i}" for i in inc] + [_SRC, "-o", out]
    subprocess.run(cmd, check=True)
    return out


def get(build_if_missing: bool = True):
    """The extension module or None (callers fall back to the numpy implementation)."""
    global _MOD
    if _MOD is not None:
        return _MOD
    path = so_path()
    if not os.path.exists(path):
        if not build_if_missing:
            return None
        try:
            build()
        except Exception:
            return None
    try:
        loader = importlib.machinery.ExtensionFileLoader("data_helpers", path)
        spec = importlib.util.spec_from_file_location("data_helpers", path, loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _MOD = mod
    except Exception:
        _MOD = None
    return _MOD
