"""Context-parallel ring attention.

Capability parity: ``nki_ring_attn_func(q, k, v, rank, src_tgt_pairs)`` as used by the reference for CP
(reference modeling_llama.py:483-484): the sequence is split contiguously over the CP group,
every rank keeps its queries and the K/V blocks travel around the ring (rank i -> i+1); each step is a blockwise
attention whose result is merged with an online-softmax (log-sum-exp) accumulator.  Causality: a block coming from
a lower rank is fully visible, the own block is causal, blocks from higher ranks are skipped.

Backward walks the ring again: dK/dV accumulators travel together with their K/V block and arrive home after CP
steps; dQ accumulates locally.  Block math uses the global LSE saved by the forward, so it is exact.
NVSwitch note: bandwidth is uniform, so ring order is only a scheduling choice (a direct all-to-all fetch is
equivalent); K/V transfers are issued before the block math so they overlap with it.

B200-native path (``_GatheredKVAttention``, taken on GPUs for bf16 / head_dim 128 / local length % 256 == 0): every GPU
has full NVLink bandwidth to every peer through NVSwitch, so the K/V shards are all-gathered ONCE (K/V of a GQA model are
small: 2 x S x Hk_local x 128 bf16) and each rank runs ONE tcgen05 flash-attention launch over its local queries against
the gathered keys with a query offset (``flash_attn_fwd(..., q_offset=rank * S_local)``: causal masking in absolute
positions, key blocks beyond the diagonal are never loaded).  The backward is one launch of the native backward with the
same offset; dK / dV come out for the whole gathered sequence and are reduce-scattered back to their owners.  Compared
with rotating blocks this removes cp - 1 kernel launches, the per-step online-softmax merges (fp32 passes over O) and
the exposed p2p latency; pulling K/V tiles from peers inside the kernel would re-read them over NVLink once per query
tile (peer memory bypasses the local L2), which is why the gather is done once up front.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.distributed as dist

from . import state as ps

_Q_CHUNK = 1024


def _native_block_ok(q, k) -> bool:
    if not q.is_cuda:
        return False
    from .. import ops
    return ops.get_backend() != "torch" and ops.C() is not None and hasattr(ops.C(), "flash_attn_fwd") \
        and q.dtype == torch.bfloat16 and q.shape[-1] == 128 and q.shape[1] == k.shape[1] and q.shape[1] % 128 == 0


def _block_fwd(q, k, v, scale: float, causal: bool):
    """q [B,Sq,H,D], k/v [B,Sk,Hk,D] -> out [B,Sq,H,D] fp32, lse [B,H,Sq] fp32."""
    if _native_block_ok(q, k):
        from .. import ops
        o, lse = ops.C().flash_attn_fwd(q.contiguous(), k.contiguous(), v.contiguous(), bool(causal), 0, float(scale))
        return o.float(), lse
    B, Sq, H, D = q.shape
    Hk = k.shape[2]
    rep = H // Hk
    kf = k.float().permute(0, 2, 3, 1)                          # [B,Hk,D,Sk]
    vf = v.float().permute(0, 2, 1, 3)                          # [B,Hk,Sk,D]
    outs, lses = [], []
    for s in range(0, Sq, _Q_CHUNK):
        e = min(Sq, s + _Q_CHUNK)
        qf = q[:, s:e].float().permute(0, 2, 1, 3).reshape(B, Hk, rep, e - s, D)
        sc = torch.matmul(qf, kf.unsqueeze(2)) * scale          # [B,Hk,rep,sq,Sk]
        if causal:
            qi = torch.arange(s, e, device=q.device).view(-1, 1)
            ki = torch.arange(k.shape[1], device=q.device).view(1, -1)
            sc = sc.masked_fill(ki > qi, float("-inf"))
        lse = torch.logsumexp(sc, dim=-1)                       # [B,Hk,rep,sq]
        p = torch.exp(sc - lse.unsqueeze(-1))
        o = torch.matmul(p, vf.unsqueeze(2))                    # [B,Hk,rep,sq,D]
        outs.append(o.reshape(B, H, e - s, D).permute(0, 2, 1, 3))
        lses.append(lse.reshape(B, H, e - s))
    return torch.cat(outs, 1), torch.cat(lses, 2)


def _block_bwd(do, q, k, v, lse, delta, scale: float, causal: bool):
    """Gradients of one block given the GLOBAL lse [B,H,Sq] and delta = rowsum(dO*O) [B,H,Sq]."""
    B, Sq, H, D = q.shape
    Hk = k.shape[2]
    rep = H // Hk
    kf = k.float().permute(0, 2, 1, 3)                          # [B,Hk,Sk,D]
    vf = v.float().permute(0, 2, 1, 3)
    dq = torch.zeros(B, Sq, H, D, dtype=torch.float32, device=q.device)
    dk = torch.zeros_like(kf)
    dv = torch.zeros_like(vf)
    for s in range(0, Sq, _Q_CHUNK):
        e = min(Sq, s + _Q_CHUNK)
        qf = q[:, s:e].float().permute(0, 2, 1, 3).reshape(B, Hk, rep, e - s, D)
        dof = do[:, s:e].float().permute(0, 2, 1, 3).reshape(B, Hk, rep, e - s, D)
        sc = torch.matmul(qf, kf.transpose(-1, -2).unsqueeze(2)) * scale
        if causal:
            qi = torch.arange(s, e, device=q.device).view(-1, 1)
            ki = torch.arange(k.shape[1], device=q.device).view(1, -1)
            sc = sc.masked_fill(ki > qi, float("-inf"))
        p = torch.exp(sc - lse[:, :, s:e].reshape(B, Hk, rep, e - s, 1))
        dv += torch.matmul(p.transpose(-1, -2), dof).sum(2)
        dp = torch.matmul(dof, vf.transpose(-1, -2).unsqueeze(2))
        ds = p * (dp - delta[:, :, s:e].reshape(B, Hk, rep, e - s, 1)) * scale
        dq[:, s:e] = torch.matmul(ds, kf.unsqueeze(2)).reshape(B, H, e - s, D).permute(0, 2, 1, 3)
        dk += torch.matmul(ds.transpose(-1, -2), qf).sum(2)
    return dq, dk.permute(0, 2, 1, 3), dv.permute(0, 2, 1, 3)


def _ring_exchange(tensors: List[torch.Tensor], group, send_to: int, recv_from: int) -> List[torch.Tensor]:
    outs = [torch.empty(t.shape, dtype=t.dtype, device=t.device) for t in tensors]
    ops = []
    for t, o in zip(tensors, outs):
        ops.append(dist.P2POp(dist.isend, t.contiguous(), send_to, group))
        ops.append(dist.P2POp(dist.irecv, o, recv_from, group))
    return outs, dist.batch_isend_irecv(ops)


class _RingAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, scale, group, ranks, my_idx):
        cp = len(ranks)
        nxt, prv = ranks[(my_idx + 1) % cp], ranks[(my_idx - 1) % cp]
        B, S, H, D = q.shape
        out = torch.zeros(B, S, H, D, dtype=torch.float32, device=q.device)
        lse = torch.full((B, H, S), float("-inf"), dtype=torch.float32, device=q.device)
        kb, vb = k, v
        for step in range(cp):
            src = (my_idx - step) % cp
            reqs = None
            if step < cp - 1:
                (kn, vn), reqs = _ring_exchange([kb, vb], group, nxt, prv)
            if src <= my_idx:
                o_s, lse_s = _block_fwd(q, kb, vb, scale, causal=(src == my_idx))
                new = torch.logaddexp(lse, lse_s)
                w_old = torch.exp(lse - new).permute(0, 2, 1).unsqueeze(-1)
                w_new = torch.exp(lse_s - new).permute(0, 2, 1).unsqueeze(-1)
                out = out * torch.nan_to_num(w_old) + o_s * w_new
                lse = new
            if reqs is not None:
                for r in reqs:
                    r.wait()
                kb, vb = kn, vn
        ctx.save_for_backward(q, k, v, out.to(q.dtype), lse)
        ctx.scale, ctx.group, ctx.ranks, ctx.my_idx = scale, group, ranks, my_idx
        return out.to(q.dtype)

    @staticmethod
    def backward(ctx, do):
        q, k, v, out, lse = ctx.saved_tensors
        scale, group, ranks, my_idx = ctx.scale, ctx.group, ctx.ranks, ctx.my_idx
        cp = len(ranks)
        nxt, prv = ranks[(my_idx + 1) % cp], ranks[(my_idx - 1) % cp]
        delta = (do.float() * out.float()).sum(-1).permute(0, 2, 1)          # [B,H,S]
        dq = torch.zeros_like(q, dtype=torch.float32)
        kb, vb = k, v
        dkb = torch.zeros_like(k, dtype=torch.float32)
        dvb = torch.zeros_like(v, dtype=torch.float32)
        for step in range(cp):
            src = (my_idx - step) % cp
            if src <= my_idx:
                dq_s, dk_s, dv_s = _block_bwd(do, q, kb, vb, lse, delta, scale, causal=(src == my_idx))
                dq += dq_s
                dkb += dk_s
                dvb += dv_s
            # pass the block (and its gradient accumulators) on; after cp hops everything is home again
            (kb, vb, dkb, dvb), reqs = _ring_exchange([kb, vb, dkb, dvb], group, nxt, prv)
            for r in reqs:
                r.wait()
        return dq.to(q.dtype), dkb.to(k.dtype), dvb.to(v.dtype), None, None, None, None


def _native_cp_ok(q: torch.Tensor, k: torch.Tensor) -> bool:
    if not q.is_cuda:
        return False
    from .. import ops
    return ops.get_backend() != "torch" and ops.C() is not None and hasattr(ops.C(), "flash_attn_bwd") \
        and q.dtype == torch.bfloat16 and q.shape[-1] == 128 and q.shape[1] == k.shape[1] and q.shape[1] % 256 == 0 \
        and q.shape[2] % k.shape[2] == 0


def _gather_seq(x: torch.Tensor, group, cp: int) -> torch.Tensor:
    """[B, S_local, H, D] on every rank -> [B, cp * S_local, H, D] in rank order."""
    B, S, H, D = x.shape
    buf = torch.empty(cp, B, S, H, D, dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(buf, x.contiguous(), group=group)
    return buf[0] if cp == 1 else (buf.view(cp * S, H, D).unsqueeze(0) if B == 1 else
                                   buf.permute(1, 0, 2, 3, 4).reshape(B, cp * S, H, D))


def _scatter_sum_seq(x: torch.Tensor, group, cp: int) -> torch.Tensor:
    """[B, cp * S_local, H, D] partial sums on every rank -> this rank's [B, S_local, H, D] slice of the total."""
    B, St, H, D = x.shape
    S = St // cp
    src = x.view(cp, S, H, D) if B == 1 else x.view(B, cp, S, H, D).permute(1, 0, 2, 3, 4).contiguous()
    out = torch.empty(B, S, H, D, dtype=x.dtype, device=x.device)
    dist.reduce_scatter_tensor(out, src.contiguous(), group=group)
    return out


class _GatheredKVAttention(torch.autograd.Function):
    """Context-parallel causal attention on the native kernels: all-gather K / V, one offset-causal launch."""

    @staticmethod
    def forward(ctx, q, k, v, scale, group, cp, my_idx):
        from .. import ops
        C = ops.C()
        kf, vf = _gather_seq(k, group, cp), _gather_seq(v, group, cp)
        q = q.contiguous()
        ops._count()
        o, lse = C.flash_attn_fwd(q, kf, vf, True, 0, float(scale), my_idx * q.shape[1])
        ctx.save_for_backward(q, kf, vf, o, lse)
        ctx.cfg = (float(scale), group, cp, my_idx)
        return o

    @staticmethod
    def backward(ctx, do):
        from .. import ops
        q, kf, vf, o, lse = ctx.saved_tensors
        scale, group, cp, my_idx = ctx.cfg
        ops._count(3)
        dq, dkf, dvf = ops.C().flash_attn_bwd(do.contiguous(), q, kf, vf, o, lse, True, 0, scale, my_idx * q.shape[1])
        # every rank produced gradients for all keys its queries could see (zeros beyond its diagonal)
        dk = _scatter_sum_seq(dkf.float(), group, cp).to(q.dtype)
        dv = _scatter_sum_seq(dvf.float(), group, cp).to(q.dtype)
        return dq, dk, dv, None, None, None, None


def ring_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: Optional[float] = None, group=None,
                   ranks: Optional[List[int]] = None) -> torch.Tensor:
    """q [B,S_local,H,D], k/v [B,S_local,Hkv,D] (this rank's contiguous sequence chunk) -> [B,S_local,H,D]."""
    group = group if group is not None else ps.get_context_model_parallel_group()
    ranks = ranks if ranks is not None else ps.get_context_model_parallel_ranks()
    scale = scale if scale is not None else 1.0 / math.sqrt(q.shape[-1])
    if len(ranks) == 1:
        from .. import ops
        return ops.flash_attention(q, k, v, causal=True, scale=scale)
    my_idx = ranks.index(dist.get_rank())
    if _native_cp_ok(q, k):
        return _GatheredKVAttention.apply(q, k, v, scale, group, len(ranks), my_idx)
    return _RingAttention.apply(q, k, v, scale, group, ranks, my_idx)
