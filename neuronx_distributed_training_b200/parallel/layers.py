"""Tensor-parallel layers: Column/Row parallel linear, GQA-aware fused QKV, vocab-parallel embedding.

Capability parity with the NxD layers the reference instantiates (reference
modeling_llama.py:185-203,298-357,550-553,730-738; transformer.py:135-178,876-955).

Design (B200-first, not a translation):
  * activations are 2-D token matrices under the hood (``[tokens, features]``); the sequence-first
    ``[S/TP, B, H]`` view is only a reshape, so every linear is exactly one GEMM call;
  * each linear is ONE autograd node that owns its collectives, which lets the same node be served
    either by the baseline path (NCCL collective + GEMM) or by the fused sm_100a kernels
    (all-gather->GEMM / GEMM->reduce-scatter over NVSwitch peer memory, ``parallel/fused.py``);
  * weight gradients are accumulated straight into an fp32 ``main_grad`` buffer when the parameter
    has one (``fp32_grad_acc``), through the GEMM epilogue on the native path.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.distributed as dist

from . import state as ps
from .grad_hooks import mark_main_grad_written
from . import mappings as mp
from .. import ops


def _default_init(std: float = 0.02) -> Callable[[torch.Tensor], None]:
    def init_(t: torch.Tensor):
        return nn.init.normal_(t, mean=0.0, std=std)
    return init_


def set_tensor_model_parallel_attributes(p: torch.Tensor, is_parallel: bool, dim: int, stride: int = 1):
    p.tensor_model_parallel = is_parallel
    p.partition_dim = dim
    p.partition_stride = stride


def param_is_tp_duplicate(p: torch.Tensor) -> bool:
    """True for a parameter that is replicated across the TP group (counted once in norms)."""
    return not getattr(p, "tensor_model_parallel", False)


def shard_master_weight(master: torch.Tensor, dim: int, tp: int, rank: int, stride: int = 1) -> torch.Tensor:
    """Slice rank's shard out of a full weight.  ``stride`` > 1 handles fused projections
    (gate||up, q||k||v): the full tensor is ``stride`` equal blocks along ``dim`` and every rank
    must hold the matching slice of each block."""
    size = master.shape[dim]
    assert size % (tp * stride) == 0, f"dim {dim} of size {size} not divisible by tp*stride={tp}*{stride}"
    if stride == 1:
        return master.chunk(tp, dim=dim)[rank].contiguous()
    blocks = master.chunk(stride, dim=dim)
    return torch.cat([b.chunk(tp, dim=dim)[rank] for b in blocks], dim=dim).contiguous()


def unshard_weights(shards, dim: int, stride: int = 1) -> torch.Tensor:
    """inverse of shard_master_weight given the list of all TP shards."""
    if stride == 1:
        return torch.cat(list(shards), dim=dim)
    per = [s.chunk(stride, dim=dim) for s in shards]
    return torch.cat([torch.cat([p[b] for p in per], dim=dim) for b in range(stride)], dim=dim)


def _init_sharded(weight: torch.Tensor, full_shape, dim: int, stride: int, init_method, cpu_init: bool,
                  group=None, seed_offset: int = 0):
    """Deterministic sharded init.  With ``cpu_init`` the full master weight is drawn on CPU from the
    current default generator and sliced, so TP-sharded init == single-device init (the reference's
    ``init_weight_cpu``, base.py:698-701).  Otherwise each rank draws its own shard on device."""
    if weight.device.type == "meta":
        return
    tp, rank = ps.group_size(group), ps.group_rank(group)
    with torch.no_grad():
        if cpu_init and tp > 1:
            master = torch.empty(full_shape, dtype=torch.float32)
            init_method(master)
            weight.copy_(shard_master_weight(master, dim, tp, rank, stride).to(weight.dtype))
        elif tp > 1 and not cpu_init:
            with _fork_rng_for_shard(weight.device):
                init_method(weight)
        else:
            init_method(weight)


class _fork_rng_for_shard:
    """TP ranks must draw different shards: use the TP-forked RNG stream when seeded."""

    def __init__(self, device):
        from .random import get_rng_tracker
        self.cm = get_rng_tracker().fork()

    def __enter__(self):
        return self.cm.__enter__()

    def __exit__(self, *a):
        return self.cm.__exit__(*a)


# ---------------------------------------------------------------------------------------------------
# weight-gradient helper
# ---------------------------------------------------------------------------------------------------
def _wgrad(dy2d: torch.Tensor, x2d: torch.Tensor, weight: torch.Tensor) -> Optional[torch.Tensor]:
    """dW = dY^T @ X.  Accumulates into ``weight.main_grad`` (fp32) when present and returns None."""
    main_grad = getattr(weight, "main_grad", None)
    if main_grad is not None:
        ops.gemm_wgrad(dy2d, x2d, out=main_grad, accumulate=True)
        mark_main_grad_written(weight)
        return None
    return ops.gemm_wgrad(dy2d, x2d).to(weight.dtype)


def _wgrad_fused_gather(fused, dy2d: torch.Tensor, x_local2d: torch.Tensor, weight: torch.Tensor):
    """dW = dY^T @ AG(X) with the all-gather of X fused into the GEMM's K loop."""
    main_grad = getattr(weight, "main_grad", None)
    if main_grad is not None:
        fused.all_gather_wgrad(dy2d, x_local2d, main_grad, True)
        mark_main_grad_written(weight)
        return None
    out = torch.empty(weight.shape, dtype=weight.dtype, device=weight.device)
    return fused.all_gather_wgrad(dy2d, x_local2d, out, False)


# Keep the all-gathered input of a sequence-parallel column linear for the weight gradient instead of gathering it a
# second time in backward (the Megatron default, which trades NVLink traffic for activation memory).  On a B200 node the
# trade goes the other way at large TP: the layer is communication-bound and 180 GB leave room for [tokens, hidden] per
# column linear.  ``model.fusions.keep_gathered_input`` / ``NXDT_KEEP_GATHERED=1``.
_KEEP_GATHERED = {"on": os.environ.get("NXDT_KEEP_GATHERED", "0") == "1"}


def set_keep_gathered_input(flag: bool) -> None:
    _KEEP_GATHERED["on"] = bool(flag)


def keep_gathered_input() -> bool:
    return _KEEP_GATHERED["on"]


class _ColumnParallelLinearFn(torch.autograd.Function):
    """Y = AG_seq(X) @ W^T (+b).  Backward: dX = RS_seq(dY @ W) (or all-reduce without SP),
    dW = AG_seq(X)^T dY."""

    @staticmethod
    def forward(ctx, x, weight, bias, group, sequence_parallel, fused, norm_weight=None, norm_eps=0.0, nvls=None):
        # ``nvls`` (parallel/nvls.py workspace) + ``norm_weight``: x is the UN-normalised sequence shard and the RMSNorm in
        # front of this linear is fused with the all-gather (one multimem.st kernel writes every rank's gathered buffer)
        ctx.group, ctx.sp, ctx.fused = group, sequence_parallel, fused
        ctx.nvls = nvls
        ctx.norm_eps = norm_eps
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        tp = ps.group_size(group)
        x2d = x.reshape(-1, x.shape[-1])
        can_fuse = fused is not None and sequence_parallel and tp > 1 and x2d.dtype == torch.bfloat16 and \
            fused.ok(x2d.shape[0], x2d.shape[1], weight.shape[0])
        ctx.can_fuse = can_fuse
        use_fused = can_fuse and fused.prefer_ag_gemm(x2d.shape[0], x2d.shape[1], weight.shape[0])
        kept = None
        keep = _KEEP_GATHERED["on"] and weight.requires_grad
        # fp8 linears (model.fusions.fp8_gemm): quantise the LOCAL rows, all-gather e4m3 bytes + one fp32 scale per row
        # (half the NVLink traffic of the bf16 gather, and no rank re-quantises rows it does not own), then the fp8 GEMM.
        # Per-row scales make this bit-identical to gather-then-quantise.  The bf16 fused kernels stand aside.
        fp8_gather = ops.fp8.enabled() and ops.fp8.recipe() == "row" and sequence_parallel and tp > 1 and x2d.dtype in (torch.bfloat16, torch.float32) \
            and os.environ.get("NXDT_FP8_GATHER", "1") != "0"
        if fp8_gather:
            xq, xs = ops.fp8.quantize_rows(x2d.contiguous())
            total_q = mp.all_gather_dim0(xq, group)
            total_s = mp.all_gather_dim0(xs, group)
            wq, ws = ops.fp8._cached(weight, "rows")
            y2d = ops.fp8.gemm(total_q, total_s, wq, ws, out_dtype=x2d.dtype)
            out_shape = (x.shape[0] * tp,) + tuple(x.shape[1:-1]) + (weight.shape[0],)
        elif sequence_parallel and tp > 1 and nvls is not None:
            x2d = x2d if x2d.is_contiguous() else x2d.contiguous()
            total_x, rstd = nvls.rmsnorm_all_gather(x2d, norm_weight, norm_eps)
            y2d = ops.gemm_nt(total_x, weight)
            if keep:
                kept = total_x.clone()           # the gathered view lives in the double-buffered NVLS workspace
            del total_x
            out_shape = (x.shape[0] * tp,) + tuple(x.shape[1:-1]) + (weight.shape[0],)
        elif sequence_parallel and tp > 1:
            if use_fused:
                if keep:
                    # the gathered view lives in the double-buffered symmetric workspace: copy it out
                    y2d, xg = fused.all_gather_gemm(x2d, weight, return_gathered=True)
                    kept = xg.clone()
                else:
                    y2d = fused.all_gather_gemm(x2d, weight)
            else:
                total_x = mp.all_gather_dim0(x2d, group)
                y2d = ops.gemm_nt(total_x, weight)
                if keep:
                    kept = total_x
                del total_x
            out_shape = (x.shape[0] * tp,) + tuple(x.shape[1:-1]) + (weight.shape[0],)
        else:
            y2d = ops.gemm_nt(x2d, weight)
            out_shape = tuple(x.shape[:-1]) + (weight.shape[0],)
        if bias is not None:
            y2d = y2d + bias
        ctx.kept = kept is not None
        if nvls is not None:
            ctx.save_for_backward(kept if kept is not None else x, weight, x2d, norm_weight, rstd)
        else:
            ctx.save_for_backward(kept if kept is not None else x, weight)
        return y2d.view(out_shape)

    @staticmethod
    def backward(ctx, dy):
        nvls = getattr(ctx, "nvls", None)
        if nvls is not None:
            x, weight, x_raw, norm_weight, rstd = ctx.saved_tensors
        else:
            x, weight = ctx.saved_tensors
        group, tp = ctx.group, ps.group_size(ctx.group)
        dy2d = dy.reshape(-1, dy.shape[-1])
        x2d = x.reshape(-1, x.shape[-1])
        need_dx = ctx.needs_input_grad[0]
        need_dw = ctx.needs_input_grad[1]
        handle = None
        total_x = x2d
        kept = getattr(ctx, "kept", False)          # x2d IS the gathered input: nothing to gather, plain wgrad GEMM
        if nvls is not None and need_dw and not kept:
            # not kept: recompute the norm and multicast it again (the counterpart of the second all-gather)
            total_x = nvls.rmsnorm_all_gather(x_raw, norm_weight, ctx.norm_eps)[0]
            kept = True
        fused_wgrad = (not kept) and ctx.can_fuse and need_dw and \
            ctx.fused.prefer_ag_wgrad(x2d.shape[0], x2d.shape[1], dy2d.shape[1])
        if ctx.sp and tp > 1 and need_dw and not kept:
            if fused_wgrad:
                pass   # the gather is fused into the K loop of the wgrad GEMM below
            else:
                # re-gather X asynchronously; it is only needed for wgrad, after dgrad
                total_x = torch.empty((x2d.shape[0] * tp, x2d.shape[1]), dtype=x2d.dtype, device=x2d.device)
                handle = dist.all_gather_into_tensor(total_x, x2d.contiguous(), group=group, async_op=True)
        dx = None
        rs_handle = None
        # A fused kernel occupies every SM and waits on its peers; an NCCL kernel in flight next to it can only be
        # tolerated if both ranks start the two in the same order, which a CUDA graph (parallel branches) does not
        # guarantee.  So an outstanding NCCL collective is completed before a fused kernel is launched.
        if need_dx:
            if ctx.sp and tp > 1:
                if ctx.can_fuse and ctx.fused.prefer_gemm_rs(dy2d.shape[0], dy2d.shape[1], weight.shape[1]):
                    if handle is not None:
                        handle.wait()
                        handle = None
                    dx = ctx.fused.gemm_reduce_scatter(dy2d, weight, trans_b=False)
                else:
                    dx_total = ops.gemm_nn(dy2d, weight)
                    dx = torch.empty((dx_total.shape[0] // tp, dx_total.shape[1]), dtype=dx_total.dtype,
                                     device=dx_total.device)
                    rs_handle = dist.reduce_scatter_tensor(dx, dx_total, group=group, async_op=True)
            else:
                dx = ops.gemm_nn(dy2d, weight)
                if tp > 1:
                    rs_handle = dist.all_reduce(dx, group=group, async_op=True)
        dw = None
        if need_dw:
            if handle is not None:
                handle.wait()
            if fused_wgrad:
                if rs_handle is not None:
                    rs_handle.wait()
                    rs_handle = None
                dw = _wgrad_fused_gather(ctx.fused, dy2d, x2d, weight)
            else:
                dw = _wgrad(dy2d, total_x, weight)
        db = dy2d.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        if rs_handle is not None:
            rs_handle.wait()
        dnw = None
        if nvls is not None and dx is not None:
            # dx is the gradient of the normalised local rows: finish with the RMSNorm backward
            dx, dnw = _rmsnorm_bwd(dx.reshape(-1, dx.shape[-1]).contiguous(), x_raw, norm_weight, rstd)
            dnw = dnw.to(norm_weight.dtype) if ctx.needs_input_grad[6] else None
        if dx is not None:
            dx = dx.view(ctx.x_shape)
        return dx, dw, db, None, None, None, dnw, None, None


class _RowParallelLinearFn(torch.autograd.Function):
    """Z = RS_seq(A @ W^T) (or all-reduce without SP).  Backward: dA = AG_seq(dZ) @ W,
    dW = AG_seq(dZ)^T @ A.

    ``pre_act="swiglu"``: the input is the fused gate||up activation and ``A = silu(gate) * up`` is computed
    inside the node.  Only gate||up is saved; A is recomputed in backward (one bandwidth-bound kernel) instead
    of being kept alive for the weight gradient -- selective recomputation of ``ActivationMultiplyMLP`` without
    a checkpoint wrapper and without re-running any GEMM."""

    @staticmethod
    def forward(ctx, x, weight, group, sequence_parallel, reduce_dtype, fused, pre_act=None):
        ctx.group, ctx.sp, ctx.fused, ctx.pre_act = group, sequence_parallel, fused, pre_act
        ctx.x_shape = x.shape
        tp = ps.group_size(group)
        x2d = x.reshape(-1, x.shape[-1])
        if pre_act == "swiglu":
            x2d = ops.swiglu(x2d)
        elif pre_act is not None:
            raise ValueError(f"unknown pre_act {pre_act!r}")
        can_fuse = fused is not None and sequence_parallel and tp > 1 and x2d.dtype == torch.bfloat16 and \
            fused.ok(x2d.shape[0] // tp, x2d.shape[1], weight.shape[0]) and x2d.shape[0] % tp == 0
        ctx.can_fuse = can_fuse
        use_fused = can_fuse and fused.prefer_gemm_rs(x2d.shape[0], x2d.shape[1], weight.shape[0])
        if tp > 1 and sequence_parallel:
            if use_fused:
                y2d = fused.gemm_reduce_scatter(x2d, weight, trans_b=True)
            else:
                y2d = mp.reduce_scatter_dim0(ops.gemm_nt(x2d, weight), group, reduce_dtype)
            out_shape = (x.shape[0] // tp,) + tuple(x.shape[1:-1]) + (weight.shape[0],)
        else:
            y2d = ops.gemm_nt(x2d, weight)
            if tp > 1:
                y2d = mp.all_reduce(y2d, group)
            out_shape = tuple(x.shape[:-1]) + (weight.shape[0],)
        ctx.save_for_backward(x, weight)
        return y2d.view(out_shape)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        group, tp = ctx.group, ps.group_size(ctx.group)
        dy2d = dy.reshape(-1, dy.shape[-1])
        x2d = x.reshape(-1, x.shape[-1])
        pre = x2d
        if ctx.pre_act == "swiglu":
            x2d = ops.swiglu(pre)                      # recompute A from the saved gate||up
        dx = dw = None
        if ctx.sp and tp > 1:
            if ctx.can_fuse and ctx.fused.prefer_ag_gemm(dy2d.shape[0], dy2d.shape[1], weight.shape[1]):
                # one fused kernel gathers dZ and multiplies; the gathered dZ is kept for wgrad
                dx, dy_total = ctx.fused.all_gather_gemm(dy2d, weight, trans_b=False, return_gathered=True)
            else:
                dy_total = mp.all_gather_dim0(dy2d, group)
                dx = ops.gemm_nn(dy_total, weight) if ctx.needs_input_grad[0] else None
        else:
            dy_total = dy2d
            dx = ops.gemm_nn(dy_total, weight) if ctx.needs_input_grad[0] else None
        if ctx.needs_input_grad[1]:
            dw = _wgrad(dy_total, x2d, weight)
        if dx is not None:
            if ctx.pre_act == "swiglu":
                del x2d
                dx = ops.swiglu_backward(dx, pre)
            dx = dx.view(ctx.x_shape)
        return dx, dw, None, None, None, None, None


# ---------------------------------------------------------------------------------------------------
# modules
# ---------------------------------------------------------------------------------------------------
def _rmsnorm_bwd(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, rstd: torch.Tensor):
    """(dx, dw fp32) of y = x * rstd * w on 2-D rows; the native kernel on CUDA, plain math elsewhere (gloo tests)."""
    if dy.is_cuda and ops.C() is not None and ops.get_backend() != "torch":
        ops._count()
        return ops.C().rmsnorm_bwd(dy, x, w, rstd)
    xhat = x.float() * rstd.float().unsqueeze(-1)
    g = dy.float() * w.float()
    dx = rstd.float().unsqueeze(-1) * (g - xhat * (g * xhat).mean(-1, keepdim=True))
    return dx.to(x.dtype), (dy.float() * xhat).sum(0)


def _nvls_for(layer, x: torch.Tensor, pre_norm) -> Optional[object]:
    """The NVLS workspace to use for ``pre_norm`` + this column linear, or None (plain norm, then the usual path)."""
    nv = getattr(layer, "nvls", None)
    if nv is None or pre_norm is None or not layer.sequence_parallel_enabled:
        return None
    w = getattr(pre_norm, "weight", None)
    if w is None or not hasattr(pre_norm, "variance_epsilon") or type(pre_norm).forward is not _plain_rmsnorm_forward():
        return None
    if not (x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.shape[-1] % 256 == 0):
        return None
    if ops.fp8.enabled() or ops.get_backend() == "torch":
        return None
    rows = x.numel() // x.shape[-1]
    if rows * nv.world * x.shape[-1] * 2 > nv.gather_bytes:
        return None
    f = getattr(layer, "fused", None)
    if f is not None:
        wt = getattr(layer, "weight", None)
        wt = wt if wt is not None else getattr(layer, "weight_qkv", None)
        # shapes where the fused all-gather -> GEMM kernel hides the transfer behind the math keep that kernel
        if wt is not None and f.ok(rows, x.shape[-1], wt.shape[0]) and f.prefer_ag_gemm(rows, x.shape[-1], wt.shape[0]):
            return None
    return nv


def _plain_rmsnorm_forward():
    from .norm import RMSNorm
    return RMSNorm.forward


class ColumnParallelLinear(nn.Module):
    """``Y = X W^T + b`` with W split along its output dimension across the TP group."""

    def __init__(self, input_size: int, output_size: int, bias: bool = True, gather_output: bool = False,
                 init_method: Optional[Callable] = None, stride: int = 1, sequence_parallel_enabled: bool = False,
                 skip_bias_add: bool = False, dtype: torch.dtype = torch.float32, device=None,
                 reduce_dtype: Optional[torch.dtype] = None, tensor_model_parallel_group=None,
                 use_cpu_initialization: bool = True):
        super().__init__()
        self.group = tensor_model_parallel_group if tensor_model_parallel_group is not None \
            else ps.get_tensor_model_parallel_group()
        tp = ps.group_size(self.group)
        assert output_size % tp == 0, f"output_size {output_size} not divisible by TP {tp}"
        self.input_size, self.output_size = input_size, output_size
        self.output_size_per_partition = output_size // tp
        self.gather_output = gather_output
        self.sequence_parallel_enabled = sequence_parallel_enabled and tp > 1
        self.skip_bias_add = skip_bias_add
        self.stride = stride
        self.reduce_dtype = reduce_dtype
        self.init_method = init_method or _default_init()
        self.use_cpu_initialization = use_cpu_initialization
        self.fused = None  # set by parallel.fused.enable_fused_tp(model)
        if self.sequence_parallel_enabled and gather_output:
            raise ValueError("sequence parallel and gather_output are mutually exclusive")
        self.weight = nn.Parameter(torch.empty(self.output_size_per_partition, input_size, dtype=dtype, device=device))
        set_tensor_model_parallel_attributes(self.weight, True, 0, stride)
        if bias:
            self.bias = nn.Parameter(torch.zeros(self.output_size_per_partition, dtype=dtype, device=device))
            set_tensor_model_parallel_attributes(self.bias, True, 0, stride)
        else:
            self.register_parameter("bias", None)
        self.init_weight_cpu()

    def init_weight_cpu(self):
        _init_sharded(self.weight, (self.output_size, self.input_size), 0, self.stride, self.init_method,
                      self.use_cpu_initialization, self.group)
        if self.bias is not None and self.bias.device.type != "meta":
            with torch.no_grad():
                self.bias.zero_()

    def forward(self, x: torch.Tensor, pre_norm: Optional[nn.Module] = None):
        """``pre_norm``: the RMSNorm module that precedes this linear.  With an NVLS workspace attached (``self.nvls``, see
        parallel/nvls.py:enable_nvls_norm_allgather) norm + sequence-parallel all-gather run as ONE multicast kernel;
        otherwise the norm is simply applied first."""
        bias = None if self.skip_bias_add else self.bias
        nv = _nvls_for(self, x, pre_norm)
        if nv is not None:
            y = _ColumnParallelLinearFn.apply(x, self.weight, bias, self.group, True, self.fused, pre_norm.weight,
                                              float(pre_norm.variance_epsilon), nv)
        else:
            if pre_norm is not None:
                x = pre_norm(x)
            y = _ColumnParallelLinearFn.apply(x, self.weight, bias, self.group, self.sequence_parallel_enabled,
                                              self.fused)
        if self.gather_output:
            y = mp.gather_from_tensor_model_parallel_region(y, self.group)
        if self.skip_bias_add:
            return y, self.bias
        return y

    def extra_repr(self):
        return (f"in={self.input_size}, out={self.output_size} (local {self.output_size_per_partition}), "
                f"sp={self.sequence_parallel_enabled}, stride={self.stride}")


class RowParallelLinear(nn.Module):
    """``Z = A W^T + b`` with W split along its input dimension across the TP group."""

    def __init__(self, input_size: int, output_size: int, bias: bool = True, input_is_parallel: bool = True,
                 init_method: Optional[Callable] = None, stride: int = 1, sequence_parallel_enabled: bool = False,
                 skip_bias_add: bool = False, dtype: torch.dtype = torch.float32, device=None,
                 reduce_dtype: Optional[torch.dtype] = None, tensor_model_parallel_group=None,
                 use_cpu_initialization: bool = True):
        super().__init__()
        self.group = tensor_model_parallel_group if tensor_model_parallel_group is not None \
            else ps.get_tensor_model_parallel_group()
        tp = ps.group_size(self.group)
        assert input_size % tp == 0, f"input_size {input_size} not divisible by TP {tp}"
        self.input_size, self.output_size = input_size, output_size
        self.input_size_per_partition = input_size // tp
        self.input_is_parallel = input_is_parallel
        self.sequence_parallel_enabled = sequence_parallel_enabled and tp > 1
        self.skip_bias_add = skip_bias_add
        self.stride = stride
        self.reduce_dtype = reduce_dtype
        self.init_method = init_method or _default_init()
        self.use_cpu_initialization = use_cpu_initialization
        self.fused = None
        self.weight = nn.Parameter(torch.empty(output_size, self.input_size_per_partition, dtype=dtype, device=device))
        set_tensor_model_parallel_attributes(self.weight, True, 1, stride)
        if bias:
            self.bias = nn.Parameter(torch.zeros(output_size, dtype=dtype, device=device))
            # bias is added after the reduction => replicated; under SP each rank sees different
            # tokens so its grad must be summed over TP like a norm weight
            self.bias.sequence_parallel_enabled = self.sequence_parallel_enabled
        else:
            self.register_parameter("bias", None)
        self.init_weight_cpu()

    def init_weight_cpu(self):
        _init_sharded(self.weight, (self.output_size, self.input_size), 1, self.stride, self.init_method,
                      self.use_cpu_initialization, self.group)
        if self.bias is not None and self.bias.device.type != "meta":
            with torch.no_grad():
                self.bias.zero_()

    def forward(self, x: torch.Tensor, pre_act: Optional[str] = None):
        if not self.input_is_parallel:
            assert pre_act is None
            x = mp.scatter_to_tensor_model_parallel_region(x, self.group)
        y = _RowParallelLinearFn.apply(x, self.weight, self.group, self.sequence_parallel_enabled,
                                       self.reduce_dtype, self.fused, pre_act)
        if self.skip_bias_add:
            return y, self.bias
        if self.bias is not None:
            y = y + self.bias
        return y

    def extra_repr(self):
        return (f"in={self.input_size} (local {self.input_size_per_partition}), out={self.output_size}, "
                f"sp={self.sequence_parallel_enabled}")


class GQAQKVColumnParallelLinear(nn.Module):
    """One module producing (q, k, v) for grouped-query attention.

    ``kv_size_multiplier`` replicates every KV head that many times (``repeat_interleave`` on the
    head axis) so the TP degree may exceed ``num_kv_heads``; the replicas on different ranks are the
    *same* logical parameter, so their gradients are summed across the replica ranks.
    ``fuse_qkv`` stores a single ``weight_qkv`` laid out ``[q_shard; k_shard; v_shard]`` per rank so
    q/k/v come out of ONE GEMM.  (reference modeling_llama.py:310-320.)
    """

    def __init__(self, input_size: int, num_attention_heads: int, num_key_value_heads: int, head_dim: int,
                 bias: bool = False, kv_size_multiplier: int = 1, fuse_qkv: bool = True,
                 init_method: Optional[Callable] = None, sequence_parallel_enabled: bool = False,
                 dtype: torch.dtype = torch.float32, device=None, tensor_model_parallel_group=None,
                 use_cpu_initialization: bool = True):
        super().__init__()
        self.group = tensor_model_parallel_group if tensor_model_parallel_group is not None \
            else ps.get_tensor_model_parallel_group()
        tp = ps.group_size(self.group)
        self.input_size = input_size
        self.num_heads, self.num_kv_heads, self.head_dim = num_attention_heads, num_key_value_heads, head_dim
        self.kv_size_multiplier = int(kv_size_multiplier)
        self.fuse_qkv = fuse_qkv
        self.sequence_parallel_enabled = sequence_parallel_enabled and tp > 1
        self.init_method = init_method or _default_init()
        self.use_cpu_initialization = use_cpu_initialization
        self.fused = None
        kv_total = num_key_value_heads * self.kv_size_multiplier
        assert num_attention_heads % tp == 0, f"num heads {num_attention_heads} % tp {tp} != 0"
        assert kv_total % tp == 0, (f"num_kv_heads*kv_size_multiplier = {kv_total} must be divisible by "
                                    f"TP {tp}; raise kv_replicator")
        self.q_out, self.kv_out = num_attention_heads * head_dim, kv_total * head_dim
        self.q_local, self.kv_local = self.q_out // tp, self.kv_out // tp
        if fuse_qkv:
            self.weight_qkv = nn.Parameter(torch.empty(self.q_local + 2 * self.kv_local, input_size,
                                                       dtype=dtype, device=device))
            set_tensor_model_parallel_attributes(self.weight_qkv, True, 0, 1)
            self.weight_qkv.fused_qkv_sizes = (self.q_local, self.kv_local, self.kv_local)
            if bias:
                self.bias_qkv = nn.Parameter(torch.zeros(self.q_local + 2 * self.kv_local, dtype=dtype, device=device))
                set_tensor_model_parallel_attributes(self.bias_qkv, True, 0, 1)
            else:
                self.register_parameter("bias_qkv", None)
        else:
            self.weight_q = nn.Parameter(torch.empty(self.q_local, input_size, dtype=dtype, device=device))
            self.weight_k = nn.Parameter(torch.empty(self.kv_local, input_size, dtype=dtype, device=device))
            self.weight_v = nn.Parameter(torch.empty(self.kv_local, input_size, dtype=dtype, device=device))
            for w in (self.weight_q, self.weight_k, self.weight_v):
                set_tensor_model_parallel_attributes(w, True, 0, 1)
            if bias:
                self.bias_q = nn.Parameter(torch.zeros(self.q_local, dtype=dtype, device=device))
                self.bias_k = nn.Parameter(torch.zeros(self.kv_local, dtype=dtype, device=device))
                self.bias_v = nn.Parameter(torch.zeros(self.kv_local, dtype=dtype, device=device))
                for b in (self.bias_q, self.bias_k, self.bias_v):
                    set_tensor_model_parallel_attributes(b, True, 0, 1)
            else:
                for n in ("bias_q", "bias_k", "bias_v"):
                    self.register_parameter(n, None)
        self._kv_replica_group = None
        if self.kv_size_multiplier > 1 and tp > 1 and dist.is_initialized():
            self._build_replica_groups(tp)
        self.initialize_weight_biases()

    # ---- kv replication ---------------------------------------------------------------------
    def replicate_kv(self, w_full: torch.Tensor) -> torch.Tensor:
        """[num_kv_heads*head_dim, in] -> [num_kv_heads*mult*head_dim, in]"""
        if self.kv_size_multiplier == 1:
            return w_full
        w = w_full.view(self.num_kv_heads, self.head_dim, -1)
        return w.repeat_interleave(self.kv_size_multiplier, dim=0).reshape(-1, w_full.shape[-1])

    def _build_replica_groups(self, tp: int):
        # ranks holding (a slice of) the same logical KV head form a replica group
        ranks = ps.get_tensor_model_parallel_ranks()
        heads_per_rank = (self.num_kv_heads * self.kv_size_multiplier) / tp
        if heads_per_rank >= self.kv_size_multiplier:
            return  # every rank holds all replicas of its heads locally: handled in-rank
        reps = int(round(self.kv_size_multiplier / max(heads_per_rank, 1)))
        # one shared set of groups per (layout, reps), created once in parallel/state.py -- not one per layer
        self._kv_replica_group, _ = ps.get_kv_replica_group(reps)
        if self._kv_replica_group is not None:
            # their gradients get one more reduction after backward: not eligible for an early data-parallel reduce
            for p in ((self.weight_qkv,) if self.fuse_qkv else (self.weight_k, self.weight_v)):
                p.kv_replica_sync = True
            # After the sync every replica holds the same K / V gradient: count those elements once in global norms.
            # ``norm_duplicate_range`` = flat [lo, hi) of this parameter that duplicates what another rank already counts
            if dist.get_rank(self._kv_replica_group) != 0:
                if self.fuse_qkv:
                    self.weight_qkv.norm_duplicate_range = (self.q_local * self.input_size, self.weight_qkv.numel())
                else:
                    for p in (self.weight_k, self.weight_v):
                        p.norm_duplicate_range = (0, p.numel())

    def _sync_replica_grads(self, *params):
        """Sum KV gradients over the ranks that hold replicas of the same logical head."""
        if self._kv_replica_group is None:
            return
        for p in params:
            g = getattr(p, "main_grad", None)
            g = g if g is not None else p.grad
            if g is not None:
                dist.all_reduce(g, group=self._kv_replica_group)

    def sync_kv_replica_grads(self):
        if self._kv_replica_group is None:
            return
        if self.fuse_qkv:
            g = getattr(self.weight_qkv, "main_grad", None)
            g = g if g is not None else self.weight_qkv.grad
            if g is not None:
                kv = g[self.q_local:]
                kvc = kv.contiguous()
                dist.all_reduce(kvc, group=self._kv_replica_group)
                kv.copy_(kvc)
        else:
            self._sync_replica_grads(self.weight_k, self.weight_v)

    # ---- init --------------------------------------------------------------------------------
    def initialize_weight_biases(self):
        tp, rank = ps.group_size(self.group), ps.group_rank(self.group)
        ws = (self.weight_qkv,) if self.fuse_qkv else (self.weight_q, self.weight_k, self.weight_v)
        if ws[0].device.type == "meta":
            return
        with torch.no_grad():
            if self.use_cpu_initialization or self.kv_size_multiplier > 1:
                q = torch.empty(self.q_out, self.input_size)
                k = torch.empty(self.num_kv_heads * self.head_dim, self.input_size)
                v = torch.empty(self.num_kv_heads * self.head_dim, self.input_size)
                for t in (q, k, v):
                    self.init_method(t)
                self.load_full_qkv(q, k, v)
            else:
                with _fork_rng_for_shard(ws[0].device):
                    for w in ws:
                        self.init_method(w)
            for n in ("bias_qkv", "bias_q", "bias_k", "bias_v"):
                b = getattr(self, n, None)
                if b is not None:
                    b.zero_()

    def load_full_qkv(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor):
        """Load un-sharded, un-replicated q/k/v weights (HF layout) into this rank's shard."""
        tp, rank = ps.group_size(self.group), ps.group_rank(self.group)
        qs = shard_master_weight(q, 0, tp, rank)
        ks = shard_master_weight(self.replicate_kv(k), 0, tp, rank)
        vs = shard_master_weight(self.replicate_kv(v), 0, tp, rank)
        with torch.no_grad():
            if self.fuse_qkv:
                self.weight_qkv.copy_(torch.cat([qs, ks, vs], 0).to(self.weight_qkv.dtype))
            else:
                self.weight_q.copy_(qs.to(self.weight_q.dtype))
                self.weight_k.copy_(ks.to(self.weight_k.dtype))
                self.weight_v.copy_(vs.to(self.weight_v.dtype))

    def forward_packed(self, x: torch.Tensor, pre_norm: Optional[nn.Module] = None) -> torch.Tensor:
        """the fused projection output ``[..., q_local + 2*kv_local]`` before it is split (fuse_qkv only)"""
        assert self.fuse_qkv
        nv = _nvls_for(self, x, pre_norm)
        if nv is not None:
            return _ColumnParallelLinearFn.apply(x, self.weight_qkv, self.bias_qkv, self.group, True, self.fused,
                                                 pre_norm.weight, float(pre_norm.variance_epsilon), nv)
        if pre_norm is not None:
            x = pre_norm(x)
        return _ColumnParallelLinearFn.apply(x, self.weight_qkv, self.bias_qkv, self.group,
                                             self.sequence_parallel_enabled, self.fused)

    def forward(self, x: torch.Tensor, pre_norm: Optional[nn.Module] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        if self.fuse_qkv:
            nv = _nvls_for(self, x, pre_norm)
            if nv is not None:
                y = _ColumnParallelLinearFn.apply(x, self.weight_qkv, self.bias_qkv, self.group, True, self.fused,
                                                  pre_norm.weight, float(pre_norm.variance_epsilon), nv)
                return torch.split(y, [self.q_local, self.kv_local, self.kv_local], dim=-1)
        if pre_norm is not None:
            x = pre_norm(x)
        if self.fuse_qkv:
            y = _ColumnParallelLinearFn.apply(x, self.weight_qkv, self.bias_qkv, self.group,
                                              self.sequence_parallel_enabled, self.fused)
            return torch.split(y, [self.q_local, self.kv_local, self.kv_local], dim=-1)
        # un-fused: gather once, three GEMMs on the gathered activations
        if self.sequence_parallel_enabled:
            x = mp.gather_from_sequence_parallel_region(x, True, self.group)
        else:
            x = mp.copy_to_tensor_model_parallel_region(x, self.group)
        q = F.linear(x, self.weight_q, self.bias_q)
        k = F.linear(x, self.weight_k, self.bias_k)
        v = F.linear(x, self.weight_v, self.bias_v)
        return q, k, v


def _native_embedding_ok(ids: torch.Tensor, weight: torch.Tensor) -> bool:
    return ids.is_cuda and ids.dim() == 2 and ids.dtype == torch.long and weight.dtype == torch.bfloat16 \
        and weight.shape[1] % 8 == 0 and ops.get_backend() != "torch" and ops.C() is not None \
        and hasattr(ops.C(), "embedding_fwd")


class _VocabParallelEmbeddingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, weight, vocab_start, vocab_end, group, sequence_parallel):
        ctx.group, ctx.sp = group, sequence_parallel
        tp = ps.group_size(group)
        ctx.native = _native_embedding_ok(ids, weight)
        ctx.num_emb = weight.shape[0]
        ctx.wdtype = weight.dtype
        if ctx.native:
            # masked lookup in one kernel, written sequence-first when the reduce-scatter follows
            ids = ids.contiguous()
            seq_first = bool(tp > 1 and sequence_parallel)
            ops._count()
            out = ops.C().embedding_fwd(ids, weight, vocab_start if tp > 1 else 0, seq_first)
            ctx.vocab_start = vocab_start if tp > 1 else 0
            ctx.weight = weight
            ctx.save_for_backward(ids)
            if tp > 1:
                out = mp.reduce_scatter_dim0(out, group) if sequence_parallel else mp.all_reduce(out, group)
            return out
        if tp > 1:
            mask = (ids < vocab_start) | (ids >= vocab_end)
            local = (ids - vocab_start).masked_fill(mask, 0)
        else:
            mask, local = None, ids
        out = F.embedding(local, weight)
        if mask is not None:
            out = out.masked_fill(mask.unsqueeze(-1), 0.0)
        ctx.save_for_backward(local, mask)
        if tp > 1:
            if sequence_parallel:
                # ids are [B, S] -> emit sequence-first [S/TP, B, H] directly
                out = mp.reduce_scatter_dim0(out.transpose(0, 1).contiguous(), group)
            else:
                out = mp.all_reduce(out, group)
        return out

    @staticmethod
    def backward(ctx, g):
        tp = ps.group_size(ctx.group)
        if ctx.native:
            (ids,) = ctx.saved_tensors
            seq_first = bool(tp > 1 and ctx.sp)
            if seq_first:
                g = mp.all_gather_dim0(g, ctx.group)                      # [S, B, H]
            g = g.contiguous()
            weight = ctx.weight
            main_grad = getattr(weight, "main_grad", None)
            if main_grad is not None and main_grad.dtype == torch.float32 and g.dtype == torch.bfloat16:
                ops._count()
                ops.C().embedding_bwd_(main_grad, ids, g, ctx.vocab_start, seq_first)
                mark_main_grad_written(weight)
                return None, None, None, None, None, None
            dw = torch.zeros(ctx.num_emb, g.shape[-1], dtype=torch.float32, device=g.device)
            ops._count()
            ops.C().embedding_bwd_(dw, ids, g.to(torch.bfloat16), ctx.vocab_start, seq_first)
            return None, dw.to(ctx.wdtype), None, None, None, None
        local, mask = ctx.saved_tensors
        if tp > 1 and ctx.sp:
            g = mp.all_gather_dim0(g, ctx.group).transpose(0, 1)
        if mask is not None:
            g = g.masked_fill(mask.unsqueeze(-1), 0.0)
        dw = torch.zeros(ctx.num_emb, g.shape[-1], dtype=torch.float32, device=g.device)
        dw.index_add_(0, local.reshape(-1), g.reshape(-1, g.shape[-1]).float())
        return None, dw.to(ctx.wdtype), None, None, None, None


class ParallelEmbedding(nn.Module):
    """Vocab-sharded embedding.  With ``sequence_parallel_enabled`` the output is emitted directly
    as ``[S/TP, B, H]`` (reduce-scatter instead of all-reduce), SURVEY §2.2 X6."""

    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: Optional[int] = None,
                 init_method: Optional[Callable] = None, dtype: torch.dtype = torch.float32, device=None,
                 sequence_parallel_enabled: bool = False, tensor_model_parallel_group=None,
                 use_cpu_initialization: bool = True):
        super().__init__()
        self.group = tensor_model_parallel_group if tensor_model_parallel_group is not None \
            else ps.get_tensor_model_parallel_group()
        tp, rank = ps.group_size(self.group), ps.group_rank(self.group)
        assert num_embeddings % tp == 0, f"vocab {num_embeddings} not divisible by TP {tp}; pad the vocab"
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.padding_idx = padding_idx
        per = num_embeddings // tp
        self.vocab_start_index, self.vocab_end_index = rank * per, (rank + 1) * per
        self.sequence_parallel_enabled = sequence_parallel_enabled and tp > 1
        self.init_method = init_method or _default_init()
        self.use_cpu_initialization = use_cpu_initialization
        self.weight = nn.Parameter(torch.empty(per, embedding_dim, dtype=dtype, device=device))
        set_tensor_model_parallel_attributes(self.weight, True, 0, 1)
        self.init_weight_cpu()

    def init_weight_cpu(self):
        _init_sharded(self.weight, (self.num_embeddings, self.embedding_dim), 0, 1, self.init_method,
                      self.use_cpu_initialization, self.group)

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        out = _VocabParallelEmbeddingFn.apply(ids, self.weight, self.vocab_start_index, self.vocab_end_index,
                                              self.group, self.sequence_parallel_enabled)
        return out
