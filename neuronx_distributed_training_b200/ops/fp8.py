"""fp8 (e4m3) GEMM path with row / column block scales (BASELINE config #4; the reference has no fp8 mode).

Recipe: every GEMM operand is quantised along its reduction dimension with ONE scale per row (a 1 x K block), the
product is dequantised by ``row_scale(m) * col_scale(n)`` in the GEMM epilogue, accumulation is fp32.  Tensors that are
reduced along their other dimension in a backward GEMM get a second, transposed quantised copy (per-original-column
scales), so forward, dgrad and wgrad all run as K-major "nt" GEMMs on ``tcgen05.mma.kind::f8f6f4``:

    fwd    Y  = q_rows(X)    . q_rows(W)^T
    dgrad  dX = q_rows(dY)   . q_colsT(W)^T
    wgrad  dW = q_colsT(dY)  . q_colsT(X)^T          (fp32 accumulate into main_grad)

The CUDA kernels (``ops/csrc/gemm_fp8_sm100.cu``) are a PREVIEW: compiled for sm_100a but not yet validated on hardware.
The PyTorch implementation below (``torch.float8_e4m3fn``) is the numerics oracle and what runs on CPU; it is exact
w.r.t. the recipe, so convergence behaviour can be tested without a GPU.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

E4M3_MAX = 448.0
# recipe: "mx"  = 1 x 32 blocks with E8M0 scales, scaled INSIDE the tensor core (tcgen05.mma.kind::mxf8f6f4.block_scale);
#         "row" = one fp32 scale per 1 x K row block, applied in the GEMM epilogue (kind::f8f6f4)
_STATE = {"enabled": os.environ.get("NXDT_FP8", "0") == "1", "native": True, "recipe": os.environ.get("NXDT_FP8_RECIPE", "mx")}


def set_recipe(name: str) -> None:
    if name not in ("mx", "row"):
        raise ValueError(f"fp8 recipe must be 'mx' or 'row', got {name!r}")
    _STATE["recipe"] = name
    clear_cache()


def recipe() -> str:
    return _STATE["recipe"]


def use_native_kernels(flag: bool) -> None:
    """Tests only: False routes quantisation and the GEMM through the PyTorch float8 reference of the same recipe."""
    _STATE["native"] = bool(flag)


def set_enabled(flag: bool) -> None:
    _STATE["enabled"] = bool(flag)


def enabled() -> bool:
    return _STATE["enabled"]


def _native(x: torch.Tensor):
    from . import C, _use_native
    if x.is_cuda and x.dtype == torch.bfloat16 and _STATE["native"] and _use_native(x) \
            and hasattr(C(), "gemm_fp8"):
        return C()
    return None


# ---- quantisation ----------------------------------------------------------------------------------------
def quantize_rows(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """x [R, C] -> (e4m3 bytes [R, C] as uint8, fp32 scale [R]);  x ~= q * scale[:, None]."""
    c = _native(x)
    if c is not None and x.is_contiguous() and x.shape[1] % 8 == 0:
        q, s = c.fp8_quantize_rows(x)
        return q, s
    xf = x.float()
    amax = xf.abs().amax(dim=1)
    scale = torch.where(amax > 0, amax / E4M3_MAX, torch.ones_like(amax))
    q = (xf / scale[:, None]).to(torch.float8_e4m3fn)
    return q.view(torch.uint8), scale


def quantize_cols_transposed(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """x [R, C] -> (e4m3 bytes of x^T [C, R], fp32 scale [C] per original column)."""
    c = _native(x)
    if c is not None and x.is_contiguous() and x.shape[1] % 2 == 0:
        q, s = c.fp8_quantize_cols_transposed(x)
        return q, s
    return quantize_rows(x.t().contiguous())


def dequantize(q: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    return q.view(torch.float8_e4m3fn).float() * scale[:, None]


# ---- GEMM ------------------------------------------------------------------------------------------------
def gemm(a_q: torch.Tensor, a_scale: torch.Tensor, b_q: torch.Tensor, b_scale: torch.Tensor,
         out: Optional[torch.Tensor] = None, accumulate: bool = False, out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """out[M, N] (+)= (a_q[M, K] b_q[N, K]^T) * a_scale[:, None] * b_scale[None, :]."""
    M, N = a_q.shape[0], b_q.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a_q.device)
        accumulate = False
    from . import C, _use_native, _count
    if a_q.is_cuda and _STATE["native"] and _use_native(a_q) and hasattr(C(), "gemm_fp8") \
            and a_q.shape[1] % 16 == 0 and N % 8 == 0 and out.is_contiguous() and out.dtype in (torch.bfloat16, torch.float32):
        _count()
        C().gemm_fp8(a_q.contiguous(), a_scale.contiguous(), b_q.contiguous(), b_scale.contiguous(), out, accumulate)
        return out
    r = (a_q.view(torch.float8_e4m3fn).float() @ b_q.view(torch.float8_e4m3fn).float().t()) * a_scale[:, None] * b_scale[None, :]
    if accumulate:
        out.add_(r.to(out.dtype))
    else:
        out.copy_(r.to(out.dtype))
    return out


# ---- the three GEMMs of a linear layer ---------------------------------------------------------------------
_WCACHE = {}


def _cached(w: torch.Tensor, kind: str):
    """Weights change once per optimizer step but are used by every micro-batch: keep their quantised copies until
    the parameter is written again (``_version`` bumps on in-place updates)."""
    key = (id(w), kind)
    hit = _WCACHE.get(key)
    ver = w._version
    if hit is not None and hit[0] == ver and hit[1] == w.data_ptr():
        return hit[2]
    if kind == "mx_rows":
        val = mx_quantize_rows(w)
    elif kind == "mx_colsT":
        val = mx_quantize_cols_transposed(w)
    else:
        val = quantize_rows(w) if kind == "rows" else quantize_cols_transposed(w)
    _WCACHE[key] = (ver, w.data_ptr(), val)
    return val


def clear_cache() -> None:
    _WCACHE.clear()


def _mx_ok(*dims: int) -> bool:
    """MX operands: rows and reduction length of both operands must be multiples of 128 (whole scale tiles)."""
    return _STATE["recipe"] == "mx" and all(d % 128 == 0 for d in dims)


def linear_fwd(x: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [M, K] @ w [N, K]^T"""
    if _mx_ok(x.shape[0], x.shape[1], w.shape[0]):
        xq, xs = mx_quantize_rows(x)
        wq, ws = _cached(w, "mx_rows")
        return gemm_mx(xq, xs, wq, ws, out=out, out_dtype=x.dtype)
    xq, xs = quantize_rows(x)
    wq, ws = _cached(w, "rows")
    return gemm(xq, xs, wq, ws, out=out, out_dtype=x.dtype)


def linear_dgrad(dy: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dy [M, N] @ w [N, K]"""
    if _mx_ok(dy.shape[0], dy.shape[1], w.shape[1]):
        gq, gs = mx_quantize_rows(dy)
        wq, ws = _cached(w, "mx_colsT")               # [K, N], blocks along N
        return gemm_mx(gq, gs, wq, ws, out=out, out_dtype=dy.dtype)
    gq, gs = quantize_rows(dy)
    wq, ws = _cached(w, "colsT")                  # [K, N], scale per input feature k
    return gemm(gq, gs, wq, ws, out=out, out_dtype=dy.dtype)


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """dy [M, N]^T @ x [M, K] -> [N, K]"""
    if _mx_ok(dy.shape[0], dy.shape[1], x.shape[1]) and (out is None or out.dtype == torch.float32 or not accumulate):
        gq, gs = mx_quantize_cols_transposed(dy)  # [N, M], blocks along the tokens
        xq, xs = mx_quantize_cols_transposed(x)   # [K, M]
        return gemm_mx(gq, gs, xq, xs, out=out, accumulate=accumulate, out_dtype=dy.dtype)
    gq, gs = quantize_cols_transposed(dy)         # [N, M]
    xq, xs = quantize_cols_transposed(x)          # [K, M]
    return gemm(gq, gs, xq, xs, out=out, accumulate=accumulate, out_dtype=dy.dtype)


# ================================================================================================================
# MX (micro-scaling) variant: one power-of-two scale (E8M0 byte) per 1 x 32 block along the reduction dimension -- the
# format `tcgen05.mma.kind::mxf8f6f4.block_scale` consumes.  Reference implementation, the scale-factor packing the tensor
# core expects, and the dispatch to the native kernel (`csrc/gemm_mxfp8_sm100.cu`: scale factors staged in tensor memory).
# ================================================================================================================
MX_BLOCK = 32


def quantize_mx(x: torch.Tensor):
    """x [R, C] (C % 32 == 0) -> (e4m3 bytes [R, C], E8M0 scale bytes [R, C/32]);  value = q * 2^(scale - 127).
    The scale is the smallest power of two that brings the block's amax into the e4m3 range (OCP MX recipe)."""
    R, C = x.shape
    assert C % MX_BLOCK == 0
    xb = x.float().reshape(R, C // MX_BLOCK, MX_BLOCK)
    amax = xb.abs().amax(dim=-1)
    # exponent e with amax / 2^e <= 448:  e = ceil(log2(amax / 448)); empty blocks get the neutral scale 2^0
    e = torch.where(amax > 0, torch.ceil(torch.log2(amax / E4M3_MAX)), torch.zeros_like(amax)).clamp(-127, 127)
    q = (xb / torch.exp2(e)[..., None]).to(torch.float8_e4m3fn)
    return q.reshape(R, C).view(torch.uint8), (e + 127).to(torch.uint8)


def _mx_native(x: torch.Tensor):
    from . import C, _use_native
    if x.is_cuda and x.dtype == torch.bfloat16 and _STATE["native"] and _use_native(x) and hasattr(C(), "mx_quantize_rows"):
        return C()
    return None


def mx_quantize_rows(x: torch.Tensor):
    """x [R, C] (C % 128 == 0) -> (e4m3 bytes [R, C], PACKED scale tiles): blocks along C.  Native kernel on the GPU."""
    c = _mx_native(x)
    if c is not None and x.is_contiguous() and x.shape[1] % 128 == 0:
        from . import _count
        _count()
        q, sf = c.mx_quantize_rows(x)
        return q, sf
    q, s = quantize_mx(x)
    return q, pack_mx_scales(s)


def mx_quantize_cols_transposed(x: torch.Tensor):
    """x [R, C] (R % 128 == 0) -> (e4m3 bytes of x^T [C, R], PACKED scale tiles): blocks along R."""
    c = _mx_native(x)
    if c is not None and x.is_contiguous() and x.shape[0] % 128 == 0 and x.shape[1] % 8 == 0:
        from . import _count
        _count()
        q, sf = c.mx_quantize_cols_transposed(x)
        return q, sf
    return mx_quantize_rows(x.t().contiguous())


def dequantize_mx(q: torch.Tensor, scale_e8m0: torch.Tensor) -> torch.Tensor:
    R, C = q.shape
    s = torch.exp2(scale_e8m0.float() - 127.0)
    return (q.view(torch.float8_e4m3fn).float().reshape(R, C // MX_BLOCK, MX_BLOCK) * s[..., None]).reshape(R, C)


def gemm_mx_reference(a_q, a_s, b_q, b_s, out_dtype=torch.bfloat16) -> torch.Tensor:
    """(A_q . B_q^T) with both operands MX-quantised along K."""
    return (dequantize_mx(a_q, a_s) @ dequantize_mx(b_q, b_s).t()).to(out_dtype)


def pack_mx_scales(scale_e8m0: torch.Tensor) -> torch.Tensor:
    """Scale bytes [R, KB] (KB = K/32 blocks) -> the tensor-core layout: tiles of 128 rows x 4 k-blocks, 512 bytes each,
    byte offset inside a tile = (r % 32) * 16 + (r // 32) * 4 + (kb % 4)   (one 16-byte line per TMEM lane, four
    32-bit words = the four 32-row groups, the byte inside a word = the k-block, selected by the MMA's scale-factor id).
    Returns [ceil(R/128), ceil(KB/4), 512] uint8 (rows / blocks beyond the matrix are the neutral scale 127)."""
    R, KB = scale_e8m0.shape
    rt, kt = (R + 127) // 128, (KB + 3) // 4
    padded = torch.full((rt * 128, kt * 4), 127, dtype=torch.uint8, device=scale_e8m0.device)
    padded[:R, :KB] = scale_e8m0
    t = padded.reshape(rt, 4, 32, kt, 4)                  # [row tile, r // 32, r % 32, k tile, kb % 4]
    return t.permute(0, 3, 2, 1, 4).reshape(rt, kt, 512).contiguous()


def unpack_mx_scales(packed: torch.Tensor, R: int, KB: int) -> torch.Tensor:
    rt, kt, _ = packed.shape
    t = packed.reshape(rt, kt, 32, 4, 4).permute(0, 3, 2, 1, 4).reshape(rt * 128, kt * 4)
    return t[:R, :KB].contiguous()


def gemm_mx(a_q: torch.Tensor, a_sf: torch.Tensor, b_q: torch.Tensor, b_sf: torch.Tensor, out: Optional[torch.Tensor] = None,
            accumulate: bool = False, out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """out[M, N] (+)= MX(a_q, a_sf) @ MX(b_q, b_sf)^T with PACKED scale tiles (``pack_mx_scales``): the block-scaled tensor-core
    kernel on the GPU (M, K multiples of 128), the dequantise-and-matmul reference elsewhere."""
    M, K = a_q.shape
    N = b_q.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a_q.device)
        accumulate = False
    from . import C, _use_native, _count
    if a_q.is_cuda and _STATE["native"] and _use_native(a_q) and hasattr(C(), "gemm_mxfp8") and M % 128 == 0 and K % 128 == 0 \
            and N % 8 == 0 and out.is_contiguous() and (out.dtype == torch.float32 or (out.dtype == torch.bfloat16 and not accumulate)):
        _count()
        C().gemm_mxfp8(a_q.contiguous(), a_sf.contiguous(), b_q.contiguous(), b_sf.contiguous(), out, accumulate)
        return out
    r = gemm_mx_reference(a_q, unpack_mx_scales(a_sf, M, K // MX_BLOCK), b_q, unpack_mx_scales(b_sf, N, K // MX_BLOCK), torch.float32)
    if accumulate:
        out.add_(r.to(out.dtype))
    else:
        out.copy_(r.to(out.dtype))
    return out
