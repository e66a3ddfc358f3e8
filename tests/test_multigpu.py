"""Multi-GPU tests (NCCL, one process per GPU).  They carry the ``gpu`` marker, skip themselves on a box with fewer GPUs
than they need, and otherwise launch ``torch.distributed.run`` on the check scripts in ``tests/gpu_checks``:

* fused AG->GEMM / GEMM->RS / AG-wgrad kernels vs fp32 matmul + exact gather,
* fused ZeRO-1 step vs the NCCL reduce-scatter + AdamW + all-gather path,
* the whole model at TP=2 (fused kernels, bf16, native flash attention) vs TP=1: loss within 1e-3 relative, every
  gradient tensor cosine >= 0.999 and norm within 2 %  (same check with the NCCL-only TP path and with the NVLS
  norm + all-gather kernel),
* the NVLS (multimem) kernels vs NCCL,
* context-parallel attention (gathered K/V + one offset-causal native launch) vs the full sequence.
Pipeline parallelism on GPUs is exercised through ``bench.py --pp`` (tests/gpu_checks/COMMANDS.md); expert parallelism,
LoRA and DPO under tensor parallelism are covered by the gloo tests (tests/test_distributed_cpu.py) and, on one GPU, by
tests/test_families_gpu.py.
"""
import os
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
pytestmark = pytest.mark.gpu
_PORT = [29350]


def _gpus() -> int:
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def run_gpus(script, nproc, ok_token, env=None, timeout=900):
    if _gpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs, this box has {_gpus()}")
    _PORT[0] += 1
    e = dict(os.environ)
    e.update({"OMP_NUM_THREADS": "1"})
    e.update(env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_PORT[0]), os.path.join(HERE, "gpu_checks", script)]
    r = subprocess.run(cmd, env=e, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0 and ok_token in r.stdout, f"{script} failed:\n{r.stdout[-4000:]}\n{r.stderr[-4000:]}"
    return r.stdout


def test_fused_tp_kernels_match_fp32_reference_2gpu():
    out = run_gpus("check_fused_tp.py", 2, "FUSED_TP_OK", {"SEQ": "4096", "FUSED_TIMELINE": "0"})
    print(out[-1500:])


def test_fused_zero1_matches_nccl_adamw_2gpu():
    run_gpus("check_fused_zero1.py", 2, "FUSED_ZERO1_OK")


@pytest.mark.parametrize("fused", ["1", "0"])
def test_whole_model_tp2_matches_tp1_on_hardware(fused):
    out = run_gpus("check_tp_model_gpu.py", 2, "TP_MODEL_GPU_OK", {"FUSED": fused})
    print(out[-1500:])


def test_nvls_multicast_kernels_2gpu():
    """multimem.st RMSNorm + all-gather and multimem.ld_reduce all-reduce (ops/csrc/nvls.cu) vs NCCL; skipped on a box whose
    symmetric allocations get no NVSwitch multicast mapping."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    try:
        out = run_gpus("check_nvls.py", 2, "NVLS_OK")
    except AssertionError as e:
        if "NVLS_UNAVAILABLE" in str(e):
            pytest.skip("no multicast mapping on this box")
        raise
    print(out[-1500:])


@pytest.mark.parametrize("keep", ["1", "0"])
def test_whole_model_tp2_with_nvls_norm_allgather_matches_tp1(keep):
    """Same whole-model parity check as above with every RMSNorm + sequence-parallel all-gather pair running as the NVLS
    kernel inside the column linear's autograd node (gathered input kept / re-multicast in backward)."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    try:
        out = run_gpus("check_tp_model_gpu.py", 2, "TP_MODEL_GPU_OK", {"FUSED": "0", "NVLS": "1", "NVLS_KEEP": keep})
    except AssertionError as e:
        if "NVLS_UNAVAILABLE" in str(e):
            pytest.skip("no multicast mapping on this box")
        raise
    print(out[-1500:])


def test_context_parallel_attention_native_kernels_2gpu():
    """CP=2: all-gathered K/V + one offset-causal tcgen05 launch per rank (forward and backward) vs the full sequence."""
    out = run_gpus("check_cp_attention.py", 2, "CP_ATTENTION_OK")
    print(out[-1200:])
