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


# ---------------------------------------------------------------------------------------------------------------------
# Memory model: does a configuration fit a 180 GB B200?  (The reference has no counterpart; on Trainium the answer comes
# from the compiler.)  Calibrated against the measured one-GPU Llama-3-8B run: 164.3 GiB peak with SwiGLU recompute.
# ---------------------------------------------------------------------------------------------------------------------
def llama_param_count(hidden_size: int, num_layers: int, vocab_size: int, intermediate_size: int, num_heads: int,
                      num_kv_heads: Optional[int] = None, head_dim: Optional[int] = None, tie_embeddings: bool = False) -> int:
    num_kv_heads = num_kv_heads or num_heads
    head_dim = head_dim or hidden_size // num_heads
    attn = hidden_size * (num_heads + 2 * num_kv_heads) * head_dim + num_heads * head_dim * hidden_size
    mlp = 3 * hidden_size * intermediate_size
    return num_layers * (attn + mlp + 2 * hidden_size) + vocab_size * hidden_size * (1 if tie_embeddings else 2) + hidden_size


def estimate_memory_gib(hidden_size: int, num_layers: int, seq_len: int, vocab_size: int, intermediate_size: int,
                        num_heads: int, num_kv_heads: Optional[int] = None, head_dim: Optional[int] = None,
                        tp: int = 1, pp: int = 1, dp: int = 1, cp: int = 1, micro_batch: int = 1,
                        recompute: Optional[str] = None, sequence_parallel: bool = True, zero1: bool = True,
                        master_weights: bool = True, fp32_grad_acc: bool = True, state_bytes: int = 4,
                        keep_gathered_input: bool = False, microbatches_in_flight: Optional[int] = None) -> dict:
    """Per-GPU memory estimate in GiB: ``{"weights", "grads", "optimizer", "activations", "workspace", "total"}``.

    ``recompute``: None | "swiglu" (ActivationMultiplyMLP) | "mlp" (LlamaMLP [+ CoreAttention, free with flash attention])
    | "full".  ``state_bytes``: bytes per element of each Adam moment (4 = fp32, 2 = bf16).  Under pipeline parallelism the
    first stage keeps ``pp`` micro-batches alive (1F1B) unless ``microbatches_in_flight`` says otherwise."""
    num_kv_heads = num_kv_heads or num_heads
    head_dim = head_dim or hidden_size // num_heads
    n_total = llama_param_count(hidden_size, num_layers, vocab_size, intermediate_size, num_heads, num_kv_heads, head_dim)
    n = n_total / (tp * pp)
    shard = dp if zero1 else 1
    weights = 2.0 * n
    grads = (4.0 if fp32_grad_acc else 2.0) * n
    optimizer = ((4.0 if master_weights else 0.0) + 2.0 * state_bytes) * n / shard
    # activations saved for backward, bytes per token and layer (bf16), flash attention (no S x S tensor)
    sp = tp if (sequence_parallel and tp > 1) else 1
    h, i = hidden_size, intermediate_size
    q, kv = num_heads * head_dim, num_kv_heads * head_dim
    per_tok = {
        "norm_inputs": 2 * h / sp,                  # the residual stream entering each of the two norms
        "col_inputs": (2 * h if keep_gathered_input and sp > 1 else 2 * h / sp),   # qkv / gate_up inputs (norm outputs)
        "qkv": (q + 2 * kv) / tp,                   # q, k, v after RoPE (attention inputs)
        "attn_out": q / tp,                         # o_proj input
        "softmax_stats": 2 * 2 * num_heads / tp,    # fp32 log-sum-exp per head, counted in bf16 units
        "gate_up": 2 * i / tp,                      # SwiGLU input
        "act": i / tp,                              # down_proj input
    }
    if recompute == "swiglu":
        per_tok["act"] = 0.0
    elif recompute == "mlp":
        per_tok["gate_up"] = per_tok["act"] = 0.0
        per_tok["col_inputs"] -= (h if keep_gathered_input and sp > 1 else h / sp)
    layers_local = num_layers / pp
    tokens = seq_len / cp * micro_batch
    if recompute == "full":
        act_layer = 2.0 * tokens * h / sp                                    # only the layer input survives
        act = act_layer * layers_local + 2.0 * tokens * sum(per_tok.values())   # + one layer being recomputed
    else:
        act = 2.0 * tokens * sum(per_tok.values()) * layers_local
    in_flight = microbatches_in_flight if microbatches_in_flight is not None else (pp if pp > 1 else 1)
    act *= in_flight
    # transient: chunked LM-head logits (fp32, 2 x 4096-token chunks), fused-TP symmetric workspace, allocator slack
    workspace = 2 * 4096 * (vocab_size / tp) * 4.0 + (2 * (tp + 2) * tokens / sp * max(h, q) * 2.0 if tp > 1 else 0.0)
    workspace += 0.01 * (weights + grads + optimizer + act)
    gib = 1024.0 ** 3
    out = {"weights": weights / gib, "grads": grads / gib, "optimizer": optimizer / gib, "activations": act / gib,
           "workspace": workspace / gib}
    out["total"] = sum(out.values())
    out["params_per_gpu_billion"] = n / 1e9
    return out


def fits_b200(estimate: dict, capacity_gib: float = 178.0) -> bool:
    """180 GB of HBM3e minus the CUDA context and NCCL buffers."""
    return estimate["total"] <= capacity_gib


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="FLOPs / memory estimate for a Llama-style configuration on B200")
    ap.add_argument("--model", default="llama3_8b", help="name under models/configs/")
    ap.add_argument("--seq-len", type=int, default=8192)
    ap.add_argument("--tp", type=int, default=1); ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--dp", type=int, default=1); ap.add_argument("--cp", type=int, default=1)
    ap.add_argument("--micro-batch", type=int, default=1)
    ap.add_argument("--recompute", default=None, choices=[None, "swiglu", "mlp", "full"])
    ap.add_argument("--state-bytes", type=int, default=4)
    ap.add_argument("--keep-gathered-input", action="store_true")
    a = ap.parse_args()
    cfg_path = os.path.join(os.path.dirname(os.path.dirname(__file__)), "models", "configs", a.model + ".json")
    c = json.load(open(cfg_path))
    dims = dict(hidden_size=c["hidden_size"], num_layers=c["num_hidden_layers"], seq_len=a.seq_len, vocab_size=c["vocab_size"],
                intermediate_size=c["intermediate_size"], num_heads=c["num_attention_heads"],
                num_kv_heads=c.get("num_key_value_heads"), head_dim=c.get("head_dim"))
    est = estimate_memory_gib(**dims, tp=a.tp, pp=a.pp, dp=a.dp, cp=a.cp, micro_batch=a.micro_batch, recompute=a.recompute,
                              state_bytes=a.state_bytes, keep_gathered_input=a.keep_gathered_input)
    fl = llama_flops_per_seq(**dims)
    print(json.dumps({"model": a.model, "flops_per_seq": fl, "flops_per_token": fl / a.seq_len,
                      "memory_gib": {k: round(v, 2) for k, v in est.items()}, "fits_b200": fits_b200(est)}, indent=1))
