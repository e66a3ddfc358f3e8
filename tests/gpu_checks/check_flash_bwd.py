"""Numerics + timing driver for the tcgen05 flash-attention backward (single GPU).

    python tests/gpu_checks/check_flash_bwd.py            # small shapes vs fp32 eager reference, then Llama-3-8B shape timing
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from neuronx_distributed_training_b200 import ops  # noqa: E402


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-9)).item()


def one(B, S, Hq, Hk, window, causal=True, fn="flash_attn_bwd"):
    C = ops.C()
    torch.manual_seed(0)
    q = torch.randn(B, S, Hq, 128, device="cuda").bfloat16().requires_grad_(True)
    k = torch.randn(B, S, Hk, 128, device="cuda").bfloat16().requires_grad_(True)
    v = torch.randn(B, S, Hk, 128, device="cuda").bfloat16().requires_grad_(True)
    do = torch.randn(B, S, Hq, 128, device="cuda").bfloat16()
    scale = 1.0 / math.sqrt(128)
    ref = ops.attention_ref(q, k, v, causal=causal, window=window or None, scale=scale)
    ref.backward(do)
    out, lse = C.flash_attn_fwd(q.detach(), k.detach(), v.detach(), causal, window, scale)
    dq, dk, dv = getattr(C, fn)(do, q.detach(), k.detach(), v.detach(), out, lse, causal, window, scale)
    torch.cuda.synchronize()
    errs = (rel(dq, q.grad), rel(dk, k.grad), rel(dv, v.grad))
    print(f"{fn} B{B} S{S} Hq{Hq} Hk{Hk} window {window} causal {causal}: rel err dq {errs[0]:.2e} dk {errs[1]:.2e} dv {errs[2]:.2e}", flush=True)
    assert all(e < 2e-2 for e in errs), errs


def timing(fn="flash_attn_bwd"):
    C = ops.C()
    bwd = getattr(C, fn)
    B, S, Hq, Hk = 1, 8192, 32, 8
    q = torch.randn(B, S, Hq, 128, device="cuda").bfloat16()
    k = torch.randn(B, S, Hk, 128, device="cuda").bfloat16()
    v = torch.randn(B, S, Hk, 128, device="cuda").bfloat16()
    do = torch.randn_like(q)
    scale = 1.0 / math.sqrt(128)
    out, lse = C.flash_attn_fwd(q, k, v, True, 0, scale)
    for _ in range(3):
        bwd(do, q, k, v, out, lse, True, 0, scale)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        bwd(do, q, k, v, out, lse, True, 0, scale)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    flops = 2.5 * 4 * B * Hq * S * S * 128 / 2          # 5 GEMMs of the backward, causal half
    print(f"{fn} backward B1 S8192 H32/8: {ms:.3f} ms, {flops / ms / 1e9:.0f} TFLOP/s", flush=True)
    qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
    o = torch.nn.functional.scaled_dot_product_attention(qq.transpose(1, 2), kk.transpose(1, 2), vv.transpose(1, 2),
                                                         is_causal=True, enable_gqa=True)
    g = do.transpose(1, 2)
    for _ in range(3):
        torch.autograd.grad(o, (qq, kk, vv), g, retain_graph=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(10):
        torch.autograd.grad(o, (qq, kk, vv), g, retain_graph=True)
    e.record()
    torch.cuda.synchronize()
    ms2 = s.elapsed_time(e) / 10
    print(f"SDPA (cuDNN) backward: {ms2:.3f} ms, {flops / ms2 / 1e9:.0f} TFLOP/s", flush=True)
    # the per-GPU shape under TP = 8 (4 query heads, 1 KV head): walks split per query head
    q8, k8, v8, do8 = q[:, :, :4].contiguous(), k[:, :, :1].contiguous(), v[:, :, :1].contiguous(), do[:, :, :4].contiguous()
    o8, l8 = C.flash_attn_fwd(q8, k8, v8, True, 0, scale)
    for _ in range(3):
        bwd(do8, q8, k8, v8, o8, l8, True, 0, scale)
    torch.cuda.synchronize()
    s.record()
    for _ in range(10):
        bwd(do8, q8, k8, v8, o8, l8, True, 0, scale)
    e.record()
    torch.cuda.synchronize()
    ms8 = s.elapsed_time(e) / 10
    qq, kk, vv = (t.clone().requires_grad_(True) for t in (q8, k8, v8))
    o = torch.nn.functional.scaled_dot_product_attention(qq.transpose(1, 2), kk.transpose(1, 2), vv.transpose(1, 2), is_causal=True, enable_gqa=True)
    g8 = do8.transpose(1, 2)
    for _ in range(3):
        torch.autograd.grad(o, (qq, kk, vv), g8, retain_graph=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(10):
        torch.autograd.grad(o, (qq, kk, vv), g8, retain_graph=True)
    e.record()
    torch.cuda.synchronize()
    ms8s = s.elapsed_time(e) / 10
    print(f"{fn} backward at TP8 (4/1 heads): {ms8:.3f} ms, {flops / 8 / ms8 / 1e9:.0f} TFLOP/s | SDPA {ms8s:.3f} ms, {flops / 8 / ms8s / 1e9:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    for cfg in [(1, 128, 1, 1, 0), (1, 256, 2, 1, 0), (2, 512, 4, 2, 0), (1, 1024, 8, 2, 0), (1, 1024, 4, 4, 300),
                (1, 2048, 4, 1, 0), (1, 2048, 2, 2, 500), (2, 4096, 16, 8, 0)]:    # the last one: one CTA per KV head (no split)
        one(*cfg)
    one(1, 512, 4, 2, 0, causal=False)
    one(2, 1024, 2, 1, 0, causal=False)
    timing()
    print("FLASH_BWD_OK")
