"""Analytic Llama FLOPs / MFU helper.

Same accounting as the reference helper (reference utils/llama_perf_estimate.py:5-134): per layer
attention (QKV, scores, context, out-proj; not discounted for causality) + three MLP GEMMs, plus the
input and output embedding GEMM-equivalents; backward = 2 x forward.  Peak numbers for B200 come from
``MEASURED_PEAKS.json`` when available.
"""
from __future__ import annotations

import json
import os
from typing import Optional

PEAK_TFLOPS_PER_NODE = {
    "trn1": 3040.0,      # 95 x 32 cores (reference constants, kept for comparison)
    "trn2": 10672.0,
    "p5": 8000.0,        # 8 x H100
    "b200x8_nominal": 8 * 2250.0,
}


def llama_flops_per_seq(hidden_size: int, num_layers: int, seq_len: int, vocab_size: int, intermediate_size: int,
                        num_heads: int, num_kv_heads: Optional[int] = None, head_dim: Optional[int] = None,
                        causal_discount: bool = False) -> float:
    """training FLOPs (fwd + bwd) for one sequence."""
    num_kv_heads = num_kv_heads or num_heads
    head_dim = head_dim or hidden_size // num_heads
    q_dim, kv_dim = num_heads * head_dim, num_kv_heads * head_dim
    qkv = 2 * seq_len * hidden_size * (q_dim + 2 * kv_dim)
    scores = 2 * seq_len * seq_len * q_dim
    ctx = 2 * seq_len * seq_len * q_dim
    if causal_discount:
        scores, ctx = scores / 2, ctx / 2
    out = 2 * seq_len * q_dim * hidden_size
    mlp = 3 * 2 * seq_len * hidden_size * intermediate_size
    emb = 2 * (2 * seq_len * hidden_size * vocab_size)
    fwd = num_layers * (qkv + scores + ctx + out + mlp) + emb
    return 3.0 * fwd


def measured_peak_tflops(default: float = 1590.0, sustained: bool = True) -> float:
    """per-GPU bf16 peak from MEASURED_PEAKS.json (fallback: the profiling guide's pool measurement)."""
    for root in (os.environ.get("GRAFT_REPO_ROOT", ""), os.getcwd(), os.path.dirname(os.path.dirname(os.path.dirname(__file__)))):
        p = os.path.join(root, "MEASURED_PEAKS.json")
        if root and os.path.exists(p):
            try:
                d = json.load(open(p))
                return float(d["bf16_tflops_sustained" if sustained and "bf16_tflops_sustained" in d else "bf16_tflops"])
            except Exception:
                pass
    return default


def calculate_mfu(seqs_per_sec: float, flops_per_seq: float, n_gpus: int, peak_tflops_per_gpu: Optional[float] = None) -> float:
    peak = (peak_tflops_per_gpu or measured_peak_tflops()) * 1e12 * n_gpus
    return seqs_per_sec * flops_per_seq / peak
