"""Sanitizer builds of the native host code ("TSAN build of the C++ host extensions").
The GPU-side counterpart (compute-sanitizer over the kernel checks) is `make sanitize-gpu`."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "neuronx_distributed_training_b200", "checkpoint", "csrc")


def _compile_and_run(tmp_path, flags, env=None):
    exe = str(tmp_path / "parallel_io_check")
    cmd = ["g++", "-std=c++17", "-O1", "-g", *flags, "-pthread", "-I", CSRC,
           os.path.join(ROOT, "tests", "native", "tsan_parallel_io.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and ("cannot find" in r.stderr or "unrecognized" in r.stderr):
        pytest.skip("sanitizer runtime not installed: " + r.stderr.strip().splitlines()[-1])
    assert r.returncode == 0, r.stderr
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([exe, str(tmp_path / "scratch.bin")], capture_output=True, text=True, env=e, timeout=300)
    return r


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_parallel_io_under_thread_sanitizer(tmp_path):
    r = _compile_and_run(tmp_path, ["-fsanitize=thread"], {"TSAN_OPTIONS": "halt_on_error=1 exitcode=66"})
    if "FATAL: ThreadSanitizer: unexpected memory mapping" in r.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow memory in this container (ASLR entropy)")
    assert r.returncode == 0 and "TSAN_PARALLEL_IO_OK" in r.stdout, r.stdout + r.stderr
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_parallel_io_under_address_and_ub_sanitizers(tmp_path):
    r = _compile_and_run(tmp_path, ["-fsanitize=address,undefined", "-fno-sanitize-recover=all"],
                         {"ASAN_OPTIONS": "detect_leaks=1"})
    assert r.returncode == 0 and "TSAN_PARALLEL_IO_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_data_helpers_under_address_sanitizer(tmp_path):
    """Builds data/csrc/helpers.cpp with ASAN + UBSAN and replays the index-builder comparison against the numpy
    oracle inside a sanitized Python process (libasan preloaded)."""
    import sysconfig
    import pybind11
    libasan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip("libasan not installed")
    so = tmp_path / ("data_helpers" + sysconfig.get_config_var("EXT_SUFFIX"))
    src = os.path.join(ROOT, "neuronx_distributed_training_b200", "data", "csrc", "helpers.cpp")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-shared", "-fPIC", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
           "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"], src, "-o", str(so)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    script = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
import data_helpers as H
rng = np.random.default_rng(0)

def sample_idx_oracle(sizes, doc_idx, seq, epochs, tpe, drop_last, extra):
    total = epochs * tpe - extra
    n = total // seq if drop_last else -(-total // seq)
    out = np.zeros((n + 1, 2), np.int64)
    di, off = 0, 0
    for s in range(1, n + 1):
        rem = seq + extra
        while rem != 0:
            if di >= len(doc_idx):
                di = len(doc_idx) - 1; off = sizes[doc_idx[di]] - extra; break
            dl = int(sizes[doc_idx[di]]) - off
            rem -= dl
            if rem <= 0:
                off += rem + dl - extra; rem = 0
            else:
                di += 1; off = 0
        out[s] = (di, off)
    return out

for trial in range(30):
    n_docs = int(rng.integers(3, 40))
    sizes = rng.integers(2, 50, size=n_docs).astype(np.int32)
    epochs = int(rng.integers(1, 4))
    doc_idx = np.concatenate([rng.permutation(n_docs) for _ in range(epochs)]).astype(np.int64)
    seq = int(rng.integers(4, 32))
    tpe = int(sizes.sum())
    for drop_last in (True, False):
        got = np.asarray(H.build_sample_idx(sizes, doc_idx, seq, epochs, tpe, drop_last, 1))
        want = sample_idx_oracle(sizes, doc_idx, seq, epochs, tpe, drop_last, 1)
        assert np.array_equal(got, want), (trial, drop_last)
    k = int(rng.integers(1, 5)); w = rng.random(k); w /= w.sum(); size = int(rng.integers(10, 300))
    di = np.zeros(size, np.uint8); dsi = np.zeros(size, np.int64)
    H.build_blending_indices(di, dsi, w.astype(np.float64), k, size, False)
    cur = np.zeros(k, np.int64)
    for i in range(size):
        d = int(np.argmax(w * max(i, 1) - cur))
        assert di[i] == d and dsi[i] == cur[d], (trial, i)
        cur[d] += 1
print("ASAN_HELPERS_OK")
"""
    e = dict(os.environ)
    e.update({"LD_PRELOAD": libasan, "ASAN_OPTIONS": "detect_leaks=0:verify_asan_link_order=0"})
    r = subprocess.run([sys.executable, "-c", script, str(tmp_path)], capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0 and "ASAN_HELPERS_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
