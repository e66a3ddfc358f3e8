"""Operator layer.

Every hot op has two implementations behind one function:
  * **native**: hand-written sm_100a CUDA kernels from ``ops/csrc`` (tcgen05/TMEM/TMA GEMM, fused
    norm / RoPE / SwiGLU / cross-entropy / AdamW, flash attention), loaded from the in-tree ``_C`` extension;
  * **torch**: plain PyTorch (cuBLAS / SDPA on GPU, CPU kernels on the gloo plumbing path).  This is the
    numerics oracle for the tests and the measured "NCCL + cuBLAS" baseline arm.

Selection: ``set_backend("native" | "torch" | "auto")``.  ``auto`` (default) uses native kernels for
CUDA tensors and torch for CPU tensors.  On a CUDA tensor with backend ``native``/``auto`` a missing
extension is an error, never a silent fallback (set ``NXDT_ALLOW_TORCH_FALLBACK=1`` to override).
"""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from . import native as _native_loader
from . import fp8 as _fp8

_BACKEND = os.environ.get("NXDT_OPS_BACKEND", "auto")
_COUNTERS = {"native_launches": 0}


def set_backend(name: str) -> None:
    global _BACKEND
    assert name in ("auto", "native", "torch")
    _BACKEND = name


def get_backend() -> str:
    return _BACKEND


def native_launch_count() -> int:
    return _COUNTERS["native_launches"]


def reset_native_launch_count() -> None:
    _COUNTERS["native_launches"] = 0


def _count(n: int = 1):
    _COUNTERS["native_launches"] += n


def _use_native(t: torch.Tensor) -> bool:
    if _BACKEND == "torch" or not t.is_cuda:
        return False
    C = _native_loader.get()
    if C is None:
        if os.environ.get("NXDT_ALLOW_TORCH_FALLBACK", "0") == "1":
            return False
        raise RuntimeError("CUDA tensor given but the sm_100a extension (ops/_C*.so) is not built/loaded: "
                           + str(_native_loader.load_error()) + "  Run `python -c 'import __graft_entry__ as g; g.build()'`.")
    return True


def C():
    return _native_loader.get()


# =====================================================================================================
# GEMM
# =====================================================================================================
_GEMM_NATIVE_DTYPES = (torch.bfloat16,)


def _gemm_native_ok(a: torch.Tensor, b: torch.Tensor, m: int, n: int, k: int) -> bool:
    if a.dtype not in _GEMM_NATIVE_DTYPES or b.dtype != a.dtype:
        return False
    # TMA needs 16-byte aligned row pitches
    return (k % 8 == 0) and (n % 8 == 0) and (m % 8 == 0) and m > 0


def _fp8_route(fp8: Optional[bool], a: torch.Tensor, b: torch.Tensor) -> bool:
    use = _fp8.enabled() if fp8 is None else fp8
    return bool(use) and a.dim() == 2 and b.dim() == 2 and a.dtype in (torch.bfloat16, torch.float32) and a.dtype == b.dtype


def gemm_nt(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, fp8: Optional[bool] = None) -> torch.Tensor:
    """out[M,N] = a[M,K] @ b[N,K]^T   (linear forward).  ``fp8``: None = follow ``ops.fp8.enabled()``."""
    if _fp8_route(fp8, a, b):
        return _fp8.linear_fwd(a, b, out)
    if _use_native(a) and _gemm_native_ok(a, b, a.shape[0], b.shape[0], a.shape[1]):
        a = a if a.is_contiguous() else a.contiguous()
        b = b if b.is_contiguous() else b.contiguous()
        if out is None:
            out = torch.empty(a.shape[0], b.shape[0], dtype=a.dtype, device=a.device)
        _count()
        C().gemm(a, b, out, True, True, False)
        return out
    r = a @ b.t()
    if out is not None:
        out.copy_(r)
        return out
    return r


def gemm_nn(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, fp8: Optional[bool] = None) -> torch.Tensor:
    """out[M,N] = a[M,K] @ b[K,N]   (dgrad: dY @ W)."""
    if _fp8_route(fp8, a, b):
        return _fp8.linear_dgrad(a, b, out)
    if _use_native(a) and _gemm_native_ok(a, b, a.shape[0], b.shape[1], a.shape[1]):
        a = a if a.is_contiguous() else a.contiguous()
        b = b if b.is_contiguous() else b.contiguous()
        if out is None:
            out = torch.empty(a.shape[0], b.shape[1], dtype=a.dtype, device=a.device)
        _count()
        C().gemm(a, b, out, True, False, False)
        return out
    r = a @ b
    if out is not None:
        out.copy_(r)
        return out
    return r


def gemm_wgrad(dy: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None,
               accumulate: bool = False, fp8: Optional[bool] = None) -> torch.Tensor:
    """out[N,K] (+)= dy[M,N]^T @ x[M,K]   (wgrad).  ``out`` may be fp32 (main_grad accumulation)."""
    if _fp8_route(fp8, dy, x):
        return _fp8.linear_wgrad(dy, x, out, accumulate)
    if _use_native(dy) and _gemm_native_ok(dy, x, dy.shape[1], x.shape[1], dy.shape[0]) and \
            (out is None or out.dtype in (torch.float32, torch.bfloat16)):
        dy = dy if dy.is_contiguous() else dy.contiguous()
        x = x if x.is_contiguous() else x.contiguous()
        if out is None:
            out = torch.empty(dy.shape[1], x.shape[1], dtype=dy.dtype, device=dy.device)
            accumulate = False
        _count()
        C().gemm(dy, x, out, False, False, accumulate)
        return out
    if out is None:
        return dy.t() @ x
    if accumulate:
        if out.dtype == dy.dtype:
            out.addmm_(dy.t(), x)
        else:
            out.add_((dy.t() @ x).to(out.dtype))
    else:
        out.copy_(dy.t() @ x)
    return out


# =====================================================================================================
# Grouped (per-expert) GEMM for dropless MoE.  Tokens are sorted by expert and every expert's run of rows is
# padded to a multiple of 128 so one persistent kernel covers all experts; the routing result never visits
# the host (reference: NKI blockwise dropless kernel).
# =====================================================================================================
GROUP_BLOCK = 128


def grouped_layout(counts: torch.Tensor, n_rows: int):
    """Padded row layout for ``n_rows`` tokens sorted by expert with ``counts[e]`` tokens each.

    Returns ``(dest, block_expert, offsets, t_cap)``: ``dest[i]`` = row of sorted token ``i`` in the padded buffer
    of ``t_cap`` rows (a static worst-case bound, so no device->host sync), ``block_expert`` int32 [t_cap/128]
    (-1 = unused block) and ``offsets`` int32 [E+1] (padded segment boundaries)."""
    E = counts.numel()
    dev = counts.device
    counts = counts.to(torch.int64)
    t_cap = (n_rows + E * (GROUP_BLOCK - 1) + GROUP_BLOCK - 1) // GROUP_BLOCK * GROUP_BLOCK
    padded = (counts + GROUP_BLOCK - 1) // GROUP_BLOCK * GROUP_BLOCK
    ends = torch.cumsum(padded, 0)
    starts = ends - padded
    ustarts = torch.cumsum(counts, 0) - counts
    e_row = torch.repeat_interleave(torch.arange(E, device=dev), counts, output_size=n_rows)
    dest = starts[e_row] + torch.arange(n_rows, device=dev) - ustarts[e_row]
    blk_start = torch.arange(t_cap // GROUP_BLOCK, device=dev) * GROUP_BLOCK
    be = torch.searchsorted(ends, blk_start, right=True)
    block_expert = torch.where(be < E, be, torch.full_like(be, -1)).to(torch.int32)
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), ends]).to(torch.int32)
    return dest, block_expert, offsets, t_cap


def _grouped_native_ok(a: torch.Tensor, w: torch.Tensor) -> bool:
    return _use_native(a) and a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and hasattr(C(), "grouped_gemm") \
        and a.shape[0] % GROUP_BLOCK == 0 and a.shape[1] % 8 == 0 and w.shape[1] % 8 == 0 and w.shape[2] % 8 == 0


def grouped_gemm(a: torch.Tensor, w: torch.Tensor, block_expert: torch.Tensor, offsets: torch.Tensor,
                 trans_w: bool = False) -> torch.Tensor:
    """Rows of expert e (``offsets[e]:offsets[e+1]``) times ``w[e]`` ([K,N]; ``trans_w``: ``w[e]`` is [N,K])."""
    n_out = w.shape[1] if trans_w else w.shape[2]
    if _grouped_native_ok(a, w):
        a = a if a.is_contiguous() else a.contiguous()
        w = w if w.is_contiguous() else w.contiguous()
        out = torch.empty(a.shape[0], n_out, dtype=a.dtype, device=a.device)
        _count()
        C().grouped_gemm(a, w, out, block_expert, trans_w)
        return out
    out = a.new_zeros(a.shape[0], n_out)
    offs = offsets.tolist()
    for e in range(w.shape[0]):
        lo, hi = offs[e], offs[e + 1]
        if hi > lo:
            out[lo:hi] = a[lo:hi] @ (w[e].t() if trans_w else w[e])
    return out


def grouped_wgrad(x: torch.Tensor, dy: torch.Tensor, offsets: torch.Tensor, out: torch.Tensor,
                  accumulate: bool) -> torch.Tensor:
    """out[e] (+)= x[rows of e]^T @ dy[rows of e];  out [E, K, N] fp32 or the activations' dtype."""
    if _use_native(x) and x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and hasattr(C(), "grouped_wgrad") \
            and x.shape[1] % GROUP_BLOCK == 0 and dy.shape[1] % 8 == 0 and x.shape[0] % GROUP_BLOCK == 0 \
            and out.dtype in (torch.float32, torch.bfloat16) and out.is_contiguous():
        if not accumulate:
            out.zero_()          # experts without tokens are skipped by the kernel
        _count()
        C().grouped_wgrad(x.contiguous(), dy.contiguous(), out, offsets, accumulate)
        return out
    offs = offsets.tolist()
    for e in range(out.shape[0]):
        lo, hi = offs[e], offs[e + 1]
        g = (x[lo:hi].t() @ dy[lo:hi]).to(out.dtype) if hi > lo else torch.zeros_like(out[e])
        if accumulate:
            out[e].add_(g)
        else:
            out[e].copy_(g)
    return out


class _GroupedLinearFn(torch.autograd.Function):
    """y[rows of e] = x[rows of e] @ w[e] with w [E, K, N]; dW accumulates into ``w.main_grad`` when present."""

    @staticmethod
    def forward(ctx, x, w, block_expert, offsets):
        ctx.save_for_backward(x, w, block_expert, offsets)
        return grouped_gemm(x, w, block_expert, offsets, trans_w=False)

    @staticmethod
    def backward(ctx, dy):
        x, w, block_expert, offsets = ctx.saved_tensors
        dy = dy.contiguous()
        dx = grouped_gemm(dy, w, block_expert, offsets, trans_w=True) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            main_grad = getattr(w, "main_grad", None)
            if main_grad is not None:
                grouped_wgrad(x, dy, offsets, main_grad, True)
                from ..parallel.grad_hooks import mark_main_grad_written
                mark_main_grad_written(w)
            else:
                dw = grouped_wgrad(x, dy, offsets, torch.empty_like(w), False)
        return dx, dw, None, None


def grouped_linear(x: torch.Tensor, w: torch.Tensor, block_expert: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    return _GroupedLinearFn.apply(x, w, block_expert, offsets)


# =====================================================================================================
# RMSNorm / LayerNorm
# =====================================================================================================
def rmsnorm_ref(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return (y * w.float()).to(x.dtype)


class _RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, eps):
        x2 = x.reshape(-1, x.shape[-1])
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        _count()
        y, rstd = C().rmsnorm_fwd(x2, w, eps)
        ctx.save_for_backward(x2, w, rstd)
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, w, rstd = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        dy2 = dy2 if dy2.is_contiguous() else dy2.contiguous()
        _count()
        dx, dw = C().rmsnorm_bwd(dy2, x2, w, rstd)
        return dx.view(ctx.shape), dw.to(w.dtype), None


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    if _use_native(x) and x.dtype in (torch.bfloat16, torch.float32) and x.shape[-1] % 8 == 0:
        return _RMSNormFn.apply(x, w, eps)
    return rmsnorm_ref(x, w, eps)


def layernorm_ref(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], eps: float = 1e-5,
                  one_plus: bool = False) -> torch.Tensor:
    ww = w + 1.0 if one_plus else w
    return F.layer_norm(x.float(), (x.shape[-1],), ww.float(), None if b is None else b.float(), eps).to(x.dtype)


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps, one_plus):
        x2 = x.reshape(-1, x.shape[-1])
        _count()
        y, mean, rstd = C().layernorm_fwd(x2, w, b, eps, one_plus)
        ctx.save_for_backward(x2, w, mean, rstd)
        ctx.one_plus, ctx.has_bias = one_plus, b is not None
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, w, mean, rstd = ctx.saved_tensors
        _count()
        dx, dw, db = C().layernorm_bwd(dy.reshape(x2.shape).contiguous(), x2, w, mean, rstd, ctx.one_plus)
        return dx.view(dy.shape), dw.to(w.dtype), (db.to(w.dtype) if ctx.has_bias else None), None, None


def layernorm(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], eps: float = 1e-5,
              one_plus: bool = False) -> torch.Tensor:
    """Affine LayerNorm (``one_plus``: Megatron layernorm1p, weight stored as w-1); fp32 statistics."""
    if _use_native(x) and hasattr(C(), "layernorm_fwd") and x.dtype in (torch.bfloat16, torch.float32) \
            and w.dtype == x.dtype and (b is None or b.dtype == x.dtype) and x.is_contiguous() \
            and x.shape[-1] % 8 == 0 and x.shape[-1] <= 8192:
        return _LayerNormFn.apply(x, w, b, eps, one_plus)
    return layernorm_ref(x, w, b, eps, one_plus)


# =====================================================================================================
# RoPE   q/k: [..., S, H, D] with cos/sin broadcastable [S, 1, D] (rotate-half convention)
# =====================================================================================================
def rotate_half(x: torch.Tensor) -> torch.Tensor:
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def rope_ref(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x [B,S,H,D]; cos/sin [S,D] (or [B,S,D]) fp32."""
    c = cos.unsqueeze(-2)
    s = sin.unsqueeze(-2)
    return (x.float() * c + rotate_half(x.float()) * s).to(x.dtype)


class _RopeFn(torch.autograd.Function):
    """In-place-capable fused RoPE on q and k together (one launch)."""

    @staticmethod
    def forward(ctx, q, k, cos, sin):
        ctx.save_for_backward(cos, sin)
        _count()
        qo, ko = C().rope_fwd(q, k, cos, sin, False)
        return qo, ko

    @staticmethod
    def backward(ctx, dq, dk):
        cos, sin = ctx.saved_tensors
        _count()
        dqo, dko = C().rope_fwd(dq.contiguous(), dk.contiguous(), cos, sin, True)
        return dqo, dko, None, None


def rope(q: torch.Tensor, k: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor):
    """q [B,S,Hq,D], k [B,S,Hk,D], cos/sin [S,D] fp32 -> rotated (q, k)."""
    if _use_native(q) and q.dtype == torch.bfloat16 and cos.dim() == 2 and q.shape[-1] % 8 == 0:
        return _RopeFn.apply(q.contiguous(), k.contiguous(), cos.contiguous(), sin.contiguous())
    return rope_ref(q, cos, sin), rope_ref(k, cos, sin)


class _QKVRopeSplitFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, packed, cos, sin, B, S, Hq, Hk):
        ctx.save_for_backward(cos, sin)
        _count()
        q, k, v = C().qkv_rope_split_fwd(packed, cos, sin, B, S, Hq, Hk)
        return q, k, v

    @staticmethod
    def backward(ctx, dq, dk, dv):
        cos, sin = ctx.saved_tensors
        _count()
        return C().qkv_rope_split_bwd(dq.contiguous(), dk.contiguous(), dv.contiguous(), cos, sin), None, None, None, None, None, None


def qkv_rope_split(packed: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, num_q_heads: int, num_kv_heads: int,
                   head_dim: int):
    """packed [S, B, (Hq+2Hk)*D] (sequence first) -> rotated q [B,S,Hq,D], k [B,S,Hk,D] and v [B,S,Hk,D], contiguous.
    One fused pass replaces split + 3x .contiguous() + RoPE + the [S,B]->[B,S] transpose."""
    S, B = packed.shape[0], packed.shape[1]
    if _use_native(packed) and packed.dtype == torch.bfloat16 and head_dim % 16 == 0 and cos.dim() == 2:
        p2 = packed.reshape(S * B, -1)
        p2 = p2 if p2.is_contiguous() else p2.contiguous()
        return _QKVRopeSplitFn.apply(p2, cos.contiguous(), sin.contiguous(), B, S, num_q_heads, num_kv_heads)
    qd, kd = num_q_heads * head_dim, num_kv_heads * head_dim
    q, k, v = torch.split(packed, [qd, kd, kd], dim=-1)
    q = q.reshape(S, B, num_q_heads, head_dim).transpose(0, 1)
    k = k.reshape(S, B, num_kv_heads, head_dim).transpose(0, 1)
    v = v.reshape(S, B, num_kv_heads, head_dim).transpose(0, 1)
    q, k = rope(q, k, cos, sin)
    return q, k, v


# =====================================================================================================
# SwiGLU on fused gate||up  ([..., 2I] -> [..., I])
# =====================================================================================================
def swiglu_ref(gate_up: torch.Tensor) -> torch.Tensor:
    g, u = gate_up.chunk(2, dim=-1)
    return F.silu(g) * u


class _SwiGLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gate_up):
        ctx.save_for_backward(gate_up)
        _count()
        return C().swiglu_fwd(gate_up)

    @staticmethod
    def backward(ctx, dout):
        (gate_up,) = ctx.saved_tensors
        _count()
        return C().swiglu_bwd(dout.contiguous(), gate_up)


def swiglu(gate_up: torch.Tensor) -> torch.Tensor:
    if _use_native(gate_up) and gate_up.dtype == torch.bfloat16 and gate_up.is_contiguous() \
            and gate_up.shape[-1] % 16 == 0:
        return _SwiGLUFn.apply(gate_up)
    return swiglu_ref(gate_up)


def swiglu_backward(dout: torch.Tensor, gate_up: torch.Tensor) -> torch.Tensor:
    """d(gate||up) of ``swiglu`` given d(out); used by nodes that recompute the activation themselves."""
    if _use_native(gate_up) and gate_up.dtype == torch.bfloat16 and gate_up.is_contiguous() \
            and gate_up.shape[-1] % 16 == 0:
        _count()
        return C().swiglu_bwd(dout.contiguous(), gate_up)
    g, u = gate_up.float().chunk(2, dim=-1)
    d = dout.float()
    sg = torch.sigmoid(g)
    return torch.cat([d * u * (sg * (1.0 + g * (1.0 - sg))), d * (g * sg)], dim=-1).to(gate_up.dtype)


_ACT_MODES = {"gelu": 0, "geglu": 1, "fast-geglu": 1, "reglu": 2, "fast-reglu": 2, "swiglu": 3, "fast-swiglu": 3,
              "relu": 4, "squared-relu": 5, "silu": 6}


def bias_act_ref(x: torch.Tensor, bias: Optional[torch.Tensor], activation: str) -> torch.Tensor:
    if bias is not None:
        x = x + bias
    if activation == "gelu":
        return F.gelu(x, approximate="tanh")
    if activation in ("swiglu", "fast-swiglu"):
        return swiglu_ref(x)
    if activation in ("geglu", "fast-geglu"):
        g, u = x.chunk(2, dim=-1)
        return F.gelu(g, approximate="tanh") * u
    if activation in ("reglu", "fast-reglu"):
        g, u = x.chunk(2, dim=-1)
        return F.relu(g) * u
    if activation == "relu":
        return F.relu(x)
    if activation == "squared-relu":
        return F.relu(x) ** 2
    if activation == "silu":
        return F.silu(x)
    raise ValueError(f"unknown activation {activation}")


class _BiasActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, mode):
        ctx.save_for_backward(x, bias)
        ctx.mode = mode
        _count()
        return C().bias_act_fwd(x, bias, mode)

    @staticmethod
    def backward(ctx, dout):
        x, bias = ctx.saved_tensors
        _count()
        dx = C().bias_act_bwd(dout.contiguous(), x, bias, ctx.mode)
        db = None
        if bias is not None and ctx.needs_input_grad[1]:
            db = dx.reshape(-1, dx.shape[-1]).sum(0, dtype=torch.float32).to(bias.dtype)
        return dx, db, None


def bias_act(x: torch.Tensor, bias: Optional[torch.Tensor], activation: str) -> torch.Tensor:
    """bias + activation for the Megatron MLP (gelu / geglu / reglu / swiglu / relu / squared-relu / silu)."""
    if activation not in _ACT_MODES:
        raise ValueError(f"unknown activation {activation}")
    mode = _ACT_MODES[activation]
    width = x.shape[-1] // 2 if mode in (1, 2, 3) else x.shape[-1]
    if _use_native(x) and hasattr(C(), "bias_act_fwd") and x.dtype in (torch.bfloat16, torch.float32) \
            and (bias is None or bias.dtype == x.dtype) and x.is_contiguous() and width % 8 == 0:
        return _BiasActFn.apply(x, bias, mode)
    return bias_act_ref(x, bias, activation)


class _BiasDropoutAddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, residual, p):
        _count()
        out, mask = C().bias_dropout_add_fwd(x, bias, residual, p)
        ctx.save_for_backward(mask)
        ctx.p, ctx.has_bias = p, bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        (mask,) = ctx.saved_tensors
        dout = dout.contiguous()
        _count()
        dx = C().dropout_bwd(dout, mask, ctx.p)
        db = dx.reshape(-1, dx.shape[-1]).sum(0, dtype=torch.float32).to(dx.dtype) if ctx.has_bias else None
        return dx, db, dout, None


def bias_dropout_add(x: torch.Tensor, bias: Optional[torch.Tensor], residual: torch.Tensor, p: float,
                     training: bool) -> torch.Tensor:
    """``residual + dropout(x + bias)`` (reference transformer.py:1453-1465); one kernel when dropout is active."""
    if p > 0.0 and training and _use_native(x) and hasattr(C(), "bias_dropout_add_fwd") \
            and x.dtype in (torch.bfloat16, torch.float32) and residual.dtype == x.dtype and x.shape == residual.shape \
            and (bias is None or bias.dtype == x.dtype) and x.is_contiguous() and residual.is_contiguous() \
            and x.shape[-1] % 8 == 0:
        return _BiasDropoutAddFn.apply(x, bias, residual, float(p))
    if bias is not None:
        x = x + bias
    if p > 0.0 and training:
        x = F.dropout(x, p=p, training=True)
    return residual + x


# =====================================================================================================
# scale + mask + softmax (eager attention path; reference: apex FusedScaleMaskSoftmax / Neuron fused softmax)
# =====================================================================================================
def scale_mask_softmax_ref(scores: torch.Tensor, scale: float, causal: bool = True, window: Optional[int] = None,
                           q_offset: int = 0) -> torch.Tensor:
    Sq, Sk = scores.shape[-2], scores.shape[-1]
    s = scores.float() * scale
    qi = torch.arange(Sq, device=scores.device).unsqueeze(-1) + q_offset
    ki = torch.arange(Sk, device=scores.device).unsqueeze(0)
    mask = torch.zeros(Sq, Sk, dtype=torch.bool, device=scores.device)
    if causal:
        mask |= ki > qi
    if window is not None and window > 0:
        mask |= ki <= qi - window
    s = s.masked_fill(mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return torch.nan_to_num(p, nan=0.0).to(scores.dtype)


class _ScaleMaskSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, scale, causal, window, q_offset):
        _count()
        y = C().scale_mask_softmax_fwd(scores, scale, causal, window, q_offset)
        ctx.save_for_backward(y)
        ctx.scale = scale
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        _count()
        return C().scale_mask_softmax_bwd(dy.contiguous(), y, ctx.scale), None, None, None, None


def scale_mask_softmax(scores: torch.Tensor, scale: float, causal: bool = True, window: Optional[int] = None,
                       q_offset: int = 0) -> torch.Tensor:
    """softmax(scale * scores + mask) over the last dim of ``[..., Sq, Sk]``; fully masked rows give zeros."""
    if _use_native(scores) and hasattr(C(), "scale_mask_softmax_fwd") and scores.is_contiguous() \
            and scores.dtype in (torch.bfloat16, torch.float32) and scores.shape[-1] % 8 == 0:
        return _ScaleMaskSoftmaxFn.apply(scores, float(scale), bool(causal), int(window or 0), int(q_offset))
    return scale_mask_softmax_ref(scores, scale, causal, window, q_offset)


def attention_eager(q, k, v, causal: bool = True, window: Optional[int] = None, scale: Optional[float] = None,
                    dropout_p: float = 0.0):
    """Materialised-scores attention in the compute dtype: batched GEMM -> fused scale-mask-softmax (-> dropout) ->
    batched GEMM.  Used where probabilities must exist (attention dropout, flash disabled); q [B,S,Hq,D]."""
    B, Sq, Hq, D = q.shape
    Hk = k.shape[2]
    scale = scale if scale is not None else 1.0 / math.sqrt(D)
    qh = q.transpose(1, 2)
    kh = k.transpose(1, 2)
    vh = v.transpose(1, 2)
    if Hk != Hq:
        kh = kh.repeat_interleave(Hq // Hk, dim=1)
        vh = vh.repeat_interleave(Hq // Hk, dim=1)
    s = torch.matmul(qh, kh.transpose(-1, -2)).contiguous()
    p = scale_mask_softmax(s, scale, causal, window, 0)
    if dropout_p > 0.0:
        p = F.dropout(p, dropout_p)
    return torch.matmul(p, vh).transpose(1, 2).contiguous()


# =====================================================================================================
# Attention   q [B,S,Hq,D], k/v [B,S,Hkv,D] -> [B,S,Hq,D]
# =====================================================================================================
def attention_ref(q, k, v, causal: bool = True, window: Optional[int] = None, scale: Optional[float] = None,
                  q_offset: int = 0, return_lse: bool = False, dropout_p: float = 0.0):
    """Eager fp32 attention (oracle).  ``q_offset`` = absolute position of q[0] relative to k[0]
    (used by ring attention blocks)."""
    B, Sq, Hq, D = q.shape
    Sk, Hk = k.shape[1], k.shape[2]
    scale = scale if scale is not None else 1.0 / math.sqrt(D)
    qf = q.float().transpose(1, 2)
    kf = k.float().transpose(1, 2).repeat_interleave(Hq // Hk, dim=1)
    vf = v.float().transpose(1, 2).repeat_interleave(Hq // Hk, dim=1)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    qi = torch.arange(Sq, device=q.device).unsqueeze(-1) + q_offset
    ki = torch.arange(Sk, device=q.device).unsqueeze(0)
    mask = torch.zeros(Sq, Sk, dtype=torch.bool, device=q.device)
    if causal:
        mask |= ki > qi
    if window is not None and window > 0:
        mask |= ki <= qi - window
    s = s.masked_fill(mask, float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    p = torch.exp(s - lse.masked_fill(lse == float("-inf"), 0.0).unsqueeze(-1))
    if dropout_p > 0.0:
        p = F.dropout(p, dropout_p)
    o = torch.matmul(p, vf).transpose(1, 2).to(q.dtype)
    return (o, lse) if return_lse else o


class _FlashAttnFn(torch.autograd.Function):
    """tcgen05 flash attention: ``flash_attn_fwd`` (two-tile ping-pong, P in TMEM) + ``flash_attn_bwd`` (transposed scores,
    P^T in TMEM, software-pipelined).  Three launches in the backward: statistics pre-pass, main kernel, fp32->bf16 dQ."""

    @staticmethod
    def forward(ctx, q, k, v, causal, window, scale):
        _count()
        o, lse = C().flash_attn_fwd(q, k, v, causal, window, scale)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.cfg = (causal, window, scale)
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        causal, window, scale = ctx.cfg
        _count(3)
        dq, dk, dv = C().flash_attn_bwd(do.contiguous(), q, k, v, o, lse, causal, window, scale)
        return dq, dk, dv, None, None, None


def _sdpa(q, k, v, causal, scale):
    # torch SDPA wants [B,H,S,D]
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                       is_causal=causal, scale=scale, enable_gqa=(q.shape[2] != k.shape[2]))
    return o.transpose(1, 2)


def flash_attention_native_ok(q, k, window=None, causal=True) -> bool:
    """Shapes the hand-written kernels cover: bf16, head_dim 128, equal q / k lengths that are a multiple of 128,
    causal (optionally with a sliding window) or full attention."""
    return (_use_native(q) and q.dtype == torch.bfloat16 and q.shape[-1] == 128 and q.shape[1] % 128 == 0
            and q.shape[1] == k.shape[1] and q.shape[2] % k.shape[2] == 0 and (causal or not window))


def flash_attention(q, k, v, causal: bool = True, window: Optional[int] = None, scale: Optional[float] = None):
    """Attention on ``[B, S, heads, D]``.  On the GPU the native kernels run forward AND backward (sliding window
    included); the library SDPA is only the ``kernels: baseline`` arm and the fallback for shapes outside the native
    coverage (head_dim != 128, ragged lengths)."""
    scale = scale if scale is not None else 1.0 / math.sqrt(q.shape[-1])
    if flash_attention_native_ok(q, k, window, causal):
        return _FlashAttnFn.apply(q.contiguous(), k.contiguous(), v.contiguous(), bool(causal), int(window or 0), float(scale))
    if q.is_cuda and (window is None or window <= 0):
        return _sdpa(q, k, v, causal, scale)
    return attention_ref(q, k, v, causal, window, scale)


# =====================================================================================================
# Cross entropy on (local) logits
# =====================================================================================================
def cross_entropy_ref(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """per-token loss [T] (0 where ignored)."""
    return F.cross_entropy(logits.float(), labels, ignore_index=ignore_index, reduction="none")


def native_ce_available(logits: torch.Tensor) -> bool:
    return (_use_native(logits) and logits.dtype == torch.bfloat16 and logits.dim() == 2 and logits.is_contiguous()
            and logits.shape[1] % 8 == 0 and hasattr(C(), "ce_stats"))


def cross_entropy_fused_(logits: torch.Tensor, labels: torch.Tensor, vocab_start: int, group, ignore_index: int = -100,
                         want_grad: bool = True):
    """Vocab-parallel CE on bf16 ``logits`` [T, V_local].  Returns (per-token loss [T] fp32, d(logits) or None);
    when ``want_grad`` the logits buffer is overwritten in place with (softmax - onehot) (0 on ignored rows)."""
    import torch.distributed as dist
    from ..parallel import state as ps
    _count(2 if want_grad else 1)
    lmax, lsum, tgt = C().ce_stats(logits, labels, int(vocab_start))
    if ps.group_size(group) > 1:
        gmax = lmax.clone()
        dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
        packed = torch.stack([lsum * torch.exp(lmax - gmax), tgt])
        dist.all_reduce(packed, group=group)
        gsum, tgt = packed[0].contiguous(), packed[1]
    else:
        gmax, gsum = lmax, lsum
    ignore = labels == ignore_index
    loss = (gsum.log() + gmax - tgt).masked_fill(ignore, 0.0)
    if not want_grad:
        return loss, None
    C().ce_grad_(logits, labels, int(vocab_start), gmax, gsum, int(ignore_index))
    return loss, logits


# =====================================================================================================
# Optimizer kernels (flat buffers)
# =====================================================================================================
def adamw_step_flat(master: Optional[torch.Tensor], param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor,
                    exp_avg_sq: torch.Tensor, lr: float, beta1: float, beta2: float, eps: float,
                    weight_decay: float, step: int, grad_scale: float = 1.0,
                    decay_mask: Optional[torch.Tensor] = None, stochastic_rounding: bool = False,
                    seed: int = 0, lr_mask: Optional[torch.Tensor] = None) -> None:
    """One fused AdamW update over a flat shard.

    ``master`` (fp32, may be None => update ``param`` in its own dtype with fp32 math), ``param`` (model
    dtype, receives the rounded copy), ``grad`` (fp32 or bf16, multiplied by ``grad_scale`` = clip coeff
    / loss scaling), fp32 moments.  ``decay_mask`` (uint8 per element) selects weight decay."""
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    if _use_native(param) and exp_avg.dtype == torch.float32:
        _count()
        if not torch.is_tensor(grad_scale):
            grad_scale = torch.full((), float(grad_scale), dtype=torch.float32, device=param.device)
        C().adamw_flat(master if master is not None else torch.empty(0, device=param.device), param, grad,
                       exp_avg, exp_avg_sq, decay_mask if decay_mask is not None else torch.empty(0, dtype=torch.uint8, device=param.device),
                       lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale.float(), stochastic_rounding, seed)
        return
    # torch path: walk the flat segment in bounded chunks so temporaries stay small (the full-segment
    # form needs ~4 fp32 temporaries of the whole shard, i.e. >100 GB for an 8B model on one GPU)
    n = param.numel()
    CH = 1 << 26
    for lo in range(0, n, CH):
        hi = min(n, lo + CH)
        _adamw_chunk_torch(None if master is None else master[lo:hi], param[lo:hi], grad[lo:hi], exp_avg[lo:hi],
                           exp_avg_sq[lo:hi], lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale,
                           None if decay_mask is None else decay_mask[lo:hi], stochastic_rounding, seed + lo)


def _adamw_chunk_torch(master, param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, bc1, bc2,
                       grad_scale, decay_mask, stochastic_rounding, seed):
    g = grad.float() * grad_scale
    p32 = master if master is not None else param.float()
    if weight_decay != 0.0:
        if decay_mask is not None:
            p32.mul_(1.0 - lr * weight_decay * decay_mask.to(p32.dtype))
        else:
            p32.mul_(1.0 - lr * weight_decay)
    m32 = exp_avg.float() if exp_avg.dtype != torch.float32 else exp_avg
    v32 = exp_avg_sq.float() if exp_avg_sq.dtype != torch.float32 else exp_avg_sq
    m32.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v32.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = (v32 / bc2).sqrt_().add_(eps)
    p32.addcdiv_(m32, denom, value=-lr / bc1)
    if m32 is not exp_avg:
        exp_avg.copy_(m32)
    if v32 is not exp_avg_sq:
        exp_avg_sq.copy_(v32)
    if master is not None:
        if stochastic_rounding and param.dtype == torch.bfloat16:
            param.copy_(stochastic_round_bf16(master, seed))
        else:
            param.copy_(master)
    elif p32 is not param:
        if stochastic_rounding and param.dtype == torch.bfloat16:
            param.copy_(stochastic_round_bf16(p32, seed))
        else:
            param.copy_(p32)


def stochastic_round_bf16(x: torch.Tensor, seed: int = 0) -> torch.Tensor:
    """fp32 -> bf16 with stochastic rounding (adds uniform noise to the 16 truncated mantissa bits)."""
    g = torch.Generator(device=x.device)
    g.manual_seed(int(seed) & 0x7FFFFFFF)
    bits = x.contiguous().view(torch.int32)
    noise = torch.randint(0, 1 << 16, bits.shape, generator=g, device=x.device, dtype=torch.int32)
    rounded = (bits + noise) & ~0xFFFF
    return rounded.view(torch.float32).to(torch.bfloat16)


def sq_norm(t: torch.Tensor) -> torch.Tensor:
    """sum of squares in fp32 (device scalar)."""
    if _use_native(t) and t.is_contiguous() and t.dtype in (torch.float32, torch.bfloat16) and t.numel() > 0:
        _count()
        return C().sq_norm(t.view(-1))
    return t.float().pow(2).sum()
