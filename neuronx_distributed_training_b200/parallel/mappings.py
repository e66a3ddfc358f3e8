"""Autograd-aware conjugate collectives for tensor / sequence parallelism.

Capability parity: NxD ``parallel_layers.mappings`` as called by the reference
(reference modeling_mixtral.py:677-679,724-726, language_model.py:320,
transformer.py:1144,1178,1193).  These are the *baseline* (NCCL / gloo) forms; the fused
sm_100a AG->GEMM / GEMM->RS kernels in ``parallel/fused.py`` replace the Column/Row hot paths.
Sequence-parallel tensors are laid out sequence-first: ``[S/TP, B, H]``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import state as ps


def _tp_group():
    return ps.get_tensor_model_parallel_group()


def _size(group) -> int:
    return ps.group_size(group)


# ---- raw collectives (no autograd) ---------------------------------------------------------------
def all_reduce(x: torch.Tensor, group=None, op=None) -> torch.Tensor:
    group = _tp_group() if group is None else group
    if _size(group) == 1:
        return x
    x = x.contiguous()
    dist.all_reduce(x, op=op or dist.ReduceOp.SUM, group=group)
    return x


def all_gather_dim0(x: torch.Tensor, group=None) -> torch.Tensor:
    group = _tp_group() if group is None else group
    n = _size(group)
    if n == 1:
        return x
    x = x.contiguous()
    out = torch.empty((x.shape[0] * n,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x, group=group)
    return out


def reduce_scatter_dim0(x: torch.Tensor, group=None, reduce_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    group = _tp_group() if group is None else group
    n = _size(group)
    if n == 1:
        return x
    assert x.shape[0] % n == 0, f"dim0 {x.shape[0]} not divisible by group size {n}"
    out_dtype = x.dtype
    if reduce_dtype is not None and reduce_dtype != x.dtype:
        x = x.to(reduce_dtype)
    x = x.contiguous()
    out = torch.empty((x.shape[0] // n,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.reduce_scatter_tensor(out, x, group=group)
    return out.to(out_dtype)


def split_dim(x: torch.Tensor, dim: int, group=None) -> torch.Tensor:
    group = _tp_group() if group is None else group
    n = _size(group)
    if n == 1:
        return x
    r = ps.group_rank(group)
    return x.chunk(n, dim=dim)[r].contiguous()


def all_gather_last_dim(x: torch.Tensor, group=None) -> torch.Tensor:
    group = _tp_group() if group is None else group
    n = _size(group)
    if n == 1:
        return x
    x = x.contiguous()
    parts = [torch.empty_like(x) for _ in range(n)]
    dist.all_gather(parts, x, group=group)
    return torch.cat(parts, dim=-1)


def split_tensor_along_last_dim(x: torch.Tensor, num_partitions: int, contiguous_split_chunks: bool = False):
    assert x.shape[-1] % num_partitions == 0
    chunks = torch.split(x, x.shape[-1] // num_partitions, dim=-1)
    return tuple(c.contiguous() for c in chunks) if contiguous_split_chunks else chunks


# ---- autograd functions --------------------------------------------------------------------------
class _CopyToTP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x

    @staticmethod
    def backward(ctx, g):
        return all_reduce(g.clone() if _size(ctx.group) > 1 else g, ctx.group), None


class _ReduceFromTP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        return all_reduce(x.clone() if _size(group) > 1 else x, group)

    @staticmethod
    def backward(ctx, g):
        return g, None


class _ScatterLastDim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return split_dim(x, -1, group)

    @staticmethod
    def backward(ctx, g):
        return all_gather_last_dim(g, ctx.group), None


class _GatherLastDim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return all_gather_last_dim(x, group)

    @staticmethod
    def backward(ctx, g):
        return split_dim(g, -1, ctx.group), None


class _ScatterSeq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return split_dim(x, 0, group)

    @staticmethod
    def backward(ctx, g):
        return all_gather_dim0(g, ctx.group), None


class _GatherSeq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group, to_model_parallel):
        ctx.group, ctx.to_mp = group, to_model_parallel
        return all_gather_dim0(x, group)

    @staticmethod
    def backward(ctx, g):
        if ctx.to_mp:
            # each TP rank computed a different function of the gathered tensor => sum the grads
            return reduce_scatter_dim0(g, ctx.group), None, None
        return split_dim(g, 0, ctx.group), None, None


class _ReduceScatterSeq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group, reduce_dtype):
        ctx.group = group
        return reduce_scatter_dim0(x, group, reduce_dtype)

    @staticmethod
    def backward(ctx, g):
        return all_gather_dim0(g, ctx.group), None, None


class _ScaleGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.scale, None


def scale_gradient(x, scale: float):
    """Identity in forward, ``grad * scale`` in backward.  Used where a computation inside a tensor-parallel region is
    REPLICATED on every rank (e.g. the MoE router): its input gradient is identical on all ranks, and the region-entry
    mapping (all-reduce / reduce-scatter in backward) would otherwise count it TP times."""
    return x if scale == 1.0 else _ScaleGrad.apply(x, scale)


def copy_to_tensor_model_parallel_region(x, group=None):
    return _CopyToTP.apply(x, _tp_group() if group is None else group)


def reduce_from_tensor_model_parallel_region(x, group=None):
    return _ReduceFromTP.apply(x, _tp_group() if group is None else group)


def scatter_to_tensor_model_parallel_region(x, group=None):
    return _ScatterLastDim.apply(x, _tp_group() if group is None else group)


def gather_from_tensor_model_parallel_region(x, group=None):
    return _GatherLastDim.apply(x, _tp_group() if group is None else group)


def scatter_to_sequence_parallel_region(x, group=None):
    return _ScatterSeq.apply(x, _tp_group() if group is None else group)


def gather_from_sequence_parallel_region(x, to_model_parallel: bool = True, group=None):
    return _GatherSeq.apply(x, _tp_group() if group is None else group, to_model_parallel)


def reduce_scatter_to_sequence_parallel_region(x, group=None, reduce_dtype=None):
    return _ReduceScatterSeq.apply(x, _tp_group() if group is None else group, reduce_dtype)
