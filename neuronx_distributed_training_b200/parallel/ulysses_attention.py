"""All-to-all context parallelism ("Ulysses"): an alternative to ring attention for ``context_parallel_size > 1``.

Each CP rank holds a contiguous chunk of the sequence for ALL of its heads.  One all-to-all turns that into the FULL
sequence for 1/CP of the heads, ordinary causal flash attention runs on it, and a second all-to-all restores the
sequence split.  The reference only offers ring attention ("Ulysses: absent ... could be offered as CP
alternative"); on an NVSwitch node every GPU reaches every peer at full bandwidth, so the all-to-all costs the same
as one hop of the ring while the attention itself stays a single library / tcgen05 flash call with perfect causal load
balance (the ring leaves rank 0 idle for CP-1 of CP steps).  Constraint: the local head counts must be divisible by
CP (KV heads are replicated up to the next multiple when they are not).

Selected by ``model.fusions.ulysses_attention: true`` (with ``ring_attention`` the default, as in the reference).
"""
from __future__ import annotations

import math
from typing import Callable

import torch
import torch.distributed as dist

from . import state as ps


def _all_to_all_equal(x: torch.Tensor, group) -> torch.Tensor:
    """x [cp, ...] -> y [cp, ...] with y[src] = what rank ``src`` had in its slot ``my_rank``."""
    cp = x.shape[0]
    out = torch.empty_like(x)
    if dist.get_backend(group) == "gloo":
        # gloo has no all_to_all: emulate with all_gather (CPU plumbing tests only)
        bufs = [torch.empty_like(x) for _ in range(cp)]
        dist.all_gather(bufs, x.contiguous(), group=group)
        me = dist.get_rank(group)
        for src in range(cp):
            out[src] = bufs[src][me]
        return out
    dist.all_to_all_single(out, x.contiguous(), group=group)
    return out


class _SeqToHead(torch.autograd.Function):
    """[B, S/cp, H, D] -> [B, S, H/cp, D]  (backward: the inverse exchange)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        cp = dist.get_world_size(group)
        B, Sl, H, D = x.shape
        send = x.reshape(B, Sl, cp, H // cp, D).permute(2, 0, 1, 3, 4).contiguous()     # [cp(dst heads), B, Sl, Hl, D]
        recv = _all_to_all_equal(send, group)                                            # [cp(src seq chunk), B, Sl, Hl, D]
        return recv.permute(1, 0, 2, 3, 4).reshape(B, cp * Sl, H // cp, D)

    @staticmethod
    def backward(ctx, g):
        return _HeadToSeq.apply(g, ctx.group), None


class _HeadToSeq(torch.autograd.Function):
    """[B, S, H/cp, D] -> [B, S/cp, H, D]."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        cp = dist.get_world_size(group)
        B, S, Hl, D = x.shape
        Sl = S // cp
        send = x.reshape(B, cp, Sl, Hl, D).permute(1, 0, 2, 3, 4).contiguous()           # [cp(dst seq chunk), B, Sl, Hl, D]
        recv = _all_to_all_equal(send, group)                                            # [cp(src heads), B, Sl, Hl, D]
        return recv.permute(1, 2, 0, 3, 4).reshape(B, Sl, cp * Hl, D)

    @staticmethod
    def backward(ctx, g):
        return _SeqToHead.apply(g, ctx.group), None


def kv_replication_for(num_kv_heads_local: int, cp: int) -> int:
    """How many times every local KV head must be repeated so that the KV head count is divisible by CP."""
    return cp // math.gcd(num_kv_heads_local, cp)


def ulysses_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, attn_fn: Callable, group=None) -> torch.Tensor:
    """q [B, S_local, Hq, D], k / v [B, S_local, Hk, D] (this rank's contiguous sequence chunk, RoPE already applied with
    the chunk's absolute positions) -> [B, S_local, Hq, D].  ``attn_fn(q, k, v)`` is a causal attention over a full
    sequence (e.g. ``CoreAttention``)."""
    group = group if group is not None else ps.get_context_model_parallel_group()
    cp = dist.get_world_size(group)
    Hq, Hk = q.shape[2], k.shape[2]
    if Hq % cp != 0:
        raise ValueError(f"ulysses attention needs the local query head count ({Hq}) to be divisible by CP ({cp})")
    rep = kv_replication_for(Hk, cp)
    if rep > 1:
        if (Hq // Hk) % rep != 0:
            raise ValueError(f"cannot balance {Hk} KV heads over CP={cp} for {Hq} query heads; use ring attention")
        k = k.repeat_interleave(rep, dim=2)
        v = v.repeat_interleave(rep, dim=2)
    qf, kf, vf = _SeqToHead.apply(q, group), _SeqToHead.apply(k, group), _SeqToHead.apply(v, group)
    of = attn_fn(qf.contiguous(), kf.contiguous(), vf.contiguous())
    return _HeadToSeq.apply(of, group)
