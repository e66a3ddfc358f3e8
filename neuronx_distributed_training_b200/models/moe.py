"""Mixture-of-experts: routers, expert MLPs, token dispatch / combine, load-balancing loss.

Capability parity: NxD ``RouterTopK`` / ``RouterSinkhorn`` / ``ExpertMLPs`` / ``MoE`` /
``load_balancing_loss_func`` as used by the reference (reference
modeling_mixtral.py:342-374,872-878, transformer.py:376-467, gpt_model.py:300-304):
  * top-k routing with softmax or sigmoid affinities (optionally re-normalised over the chosen k), or Sinkhorn;
  * experts stored stacked (``[E, ...]``), tensor-parallel sharded on the FFN dimension (gate/up column, down row);
  * capacity-factor token dropping, or **dropless** (capacity_factor = 0): tokens are sorted by expert and every
    expert runs on its exact contiguous slice (Megablocks-style), no padding, no drops;
  * expert parallelism: experts split over the EP group, tokens exchanged with all-to-all (NVSwitch gives every
    pair full bandwidth, so a direct all-to-all is the right primitive; gloo falls back to all-gather);
  * sequence-parallel aware: the block gathers the sequence before routing and reduce-scatters after.
"""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..parallel import state as ps
from ..parallel import mappings as mp
from ..parallel.layers import set_tensor_model_parallel_attributes, shard_master_weight


def load_balancing_loss_func(router_logits, num_experts: int, top_k: int) -> torch.Tensor:
    """Switch-transformer auxiliary loss: E * sum_e f_e * P_e, with f the fraction of (token, slot) assignments and
    P the mean router probability.  ``router_logits``: tensor [T, E] or a tuple/list of them (one per MoE layer)."""
    if isinstance(router_logits, (tuple, list)):
        router_logits = torch.cat([r.reshape(-1, num_experts) for r in router_logits if r is not None], dim=0)
    probs = F.softmax(router_logits.float(), dim=-1)
    _, sel = torch.topk(probs, top_k, dim=-1)
    mask = F.one_hot(sel, num_experts).float()                 # [T, k, E]
    tokens_per_expert = mask.mean(dim=0)                       # [k, E]
    router_prob_per_expert = probs.mean(dim=0)                 # [E]
    return (tokens_per_expert * router_prob_per_expert.unsqueeze(0)).sum() * num_experts


# how often the grouped (one-launch-for-all-experts) tcgen05 path ran: lets tests assert it was taken
_GROUPED_CALLS = {"n": 0}


class RouterBase(nn.Module):
    def __init__(self, num_experts: int, top_k: int, hidden_size: int, act_fn: str = "softmax", dtype=torch.float32,
                 device=None, sequence_parallel_enabled: bool = False):
        super().__init__()
        self.num_experts, self.top_k, self.act_fn = num_experts, top_k, act_fn
        # the router is replicated over TP; under SP its grad is a partial sum over tokens
        self.linear_router = nn.Linear(hidden_size, num_experts, bias=False, dtype=dtype, device=device)
        nn.init.normal_(self.linear_router.weight, 0.0, 0.02)
        self.linear_router.weight.sequence_parallel_enabled = False

    def logits(self, x: torch.Tensor) -> torch.Tensor:
        return F.linear(x.float(), self.linear_router.weight.float())

    def affinities(self, logits: torch.Tensor) -> torch.Tensor:
        if self.act_fn == "sigmoid":
            return torch.sigmoid(logits)
        return F.softmax(logits, dim=-1)


class RouterTopK(RouterBase):
    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """x [T,H] -> (router_logits [T,E], affinities [T,E], expert_index [T,k])"""
        lg = self.logits(x)
        aff = self.affinities(lg)
        _, idx = torch.topk(lg, self.top_k, dim=-1)
        return lg, aff, idx


class RouterSinkhorn(RouterBase):
    """Top-1 routing on a Sinkhorn-normalised assignment (balanced without an aux loss); affinities for the
    combine step come from the activation of the raw logits."""

    def __init__(self, *a, sinkhorn_iterations: int = 30, sinkhorn_tol: Optional[float] = None, **kw):
        super().__init__(*a, **kw)
        self.iters, self.tol = sinkhorn_iterations, sinkhorn_tol

    @staticmethod
    def sinkhorn(cost: torch.Tensor, iters: int, tol: Optional[float]) -> torch.Tensor:
        cost = torch.exp(cost.float())
        d0 = torch.ones(cost.shape[0], device=cost.device)
        d1 = torch.ones(cost.shape[1], device=cost.device)
        eps = 1e-8
        for _ in range(iters):
            d0_new = (1.0 / d0.numel()) / ((d1.unsqueeze(0) * cost).sum(1) + eps)
            d1 = (1.0 / d1.numel()) / ((d0_new.unsqueeze(1) * cost).sum(0) + eps)
            if tol is not None and (d0_new - d0).abs().mean() < tol:
                d0 = d0_new
                break
            d0 = d0_new
        return d1.unsqueeze(0) * cost * d0.unsqueeze(1)

    def forward(self, x):
        lg = self.logits(x)
        aff = self.affinities(lg)
        if self.training:
            with torch.no_grad():
                norm = self.sinkhorn(lg.detach(), self.iters, self.tol)
                _, idx = torch.topk(norm, self.top_k, dim=-1)
        else:
            _, idx = torch.topk(lg, self.top_k, dim=-1)
        return lg, aff, idx


class ExpertMLPs(nn.Module):
    """Stacked expert FFNs.  ``gate_up_proj`` [E_local, H, 2*I/tp] (or [E,H,I/tp] without GLU), ``down_proj``
    [E_local, I/tp, H]; names follow the reference converter's ``expert_mlps.mlp_op.{gate_up,down}_proj``."""

    def __init__(self, num_experts: int, hidden_size: int, intermediate_size: int, glu: bool = True,
                 activation: str = "silu", dtype=torch.bfloat16, device=None, init_std: float = 0.02,
                 use_cpu_initialization: bool = False):
        super().__init__()
        tp, tpr = ps.get_tensor_model_parallel_size(), ps.get_tensor_model_parallel_rank()
        ep, epr = ps.get_expert_model_parallel_size(), ps.get_expert_model_parallel_rank()
        assert num_experts % ep == 0 and intermediate_size % tp == 0
        self.E, self.E_local, self.ep, self.ep_rank = num_experts, num_experts // ep, ep, epr
        self.H, self.I_local, self.glu, self.activation = hidden_size, intermediate_size // tp, glu, activation
        mult = 2 if glu else 1
        self.mlp_op = nn.Module()
        self.mlp_op.gate_up_proj = nn.Parameter(torch.empty(self.E_local, hidden_size, mult * self.I_local, dtype=dtype, device=device))
        self.mlp_op.down_proj = nn.Parameter(torch.empty(self.E_local, self.I_local, hidden_size, dtype=dtype, device=device))
        set_tensor_model_parallel_attributes(self.mlp_op.gate_up_proj, True, 2, mult)
        set_tensor_model_parallel_attributes(self.mlp_op.down_proj, True, 1, 1)
        for p in (self.mlp_op.gate_up_proj, self.mlp_op.down_proj):
            p.expert_parallel = ep > 1
        with torch.no_grad():
            if use_cpu_initialization and (tp > 1 or ep > 1):
                g = torch.empty(num_experts, hidden_size, mult * intermediate_size).normal_(0, init_std)
                d = torch.empty(num_experts, intermediate_size, hidden_size).normal_(0, init_std)
                g = g[epr * self.E_local:(epr + 1) * self.E_local]
                d = d[epr * self.E_local:(epr + 1) * self.E_local]
                self.mlp_op.gate_up_proj.copy_(shard_master_weight(g, 2, tp, tpr, mult))
                self.mlp_op.down_proj.copy_(shard_master_weight(d, 1, tp, tpr))
            else:
                self.mlp_op.gate_up_proj.normal_(0, init_std)
                self.mlp_op.down_proj.normal_(0, init_std)

    def _act(self, h: torch.Tensor) -> torch.Tensor:
        if self.glu:
            if self.activation in ("silu", "swiglu"):
                return ops.swiglu(h.contiguous()) if h.dim() == 2 else ops.swiglu_ref(h)
            g, u = h.chunk(2, dim=-1)
            return ops.bias_act(g, None, self.activation) * u
        return ops.bias_act(h, None, self.activation)

    def forward_grouped(self, x_sorted: torch.Tensor, tokens_per_expert) -> torch.Tensor:
        """Dropless expert FFNs as two grouped GEMMs (+ one for each gradient): tokens are scattered into a layout
        where every expert's rows are padded to a multiple of 128, built on the device from the routing counts,
        so there is no host sync and no per-expert launch (the reference's blockwise NKI kernel, X19)."""
        dev = x_sorted.device
        counts = tokens_per_expert if torch.is_tensor(tokens_per_expert) else torch.tensor(list(tokens_per_expert))
        counts = counts.to(dev)
        _GROUPED_CALLS["n"] += 1
        n = x_sorted.shape[0]
        dest, block_expert, offsets, t_cap = ops.grouped_layout(counts, n)
        xp = x_sorted.new_zeros(t_cap, x_sorted.shape[1]).index_copy(0, dest, x_sorted)
        h = ops.grouped_linear(xp, self.mlp_op.gate_up_proj, block_expert, offsets)
        h = self._act(h)
        y = ops.grouped_linear(h, self.mlp_op.down_proj, block_expert, offsets)
        return y.index_select(0, dest)

    def _use_grouped(self, x: torch.Tensor) -> bool:
        mode = os.environ.get("NXDT_MOE_GROUPED", "auto")
        if mode in ("0", "off"):
            return False
        if mode in ("1", "on"):
            return True
        return x.is_cuda and x.dtype == torch.bfloat16 and ops.get_backend() != "torch" and ops.C() is not None \
            and hasattr(ops.C(), "grouped_gemm") and self.H % 128 == 0 and self.I_local % 128 == 0

    def forward_sorted(self, x_sorted: torch.Tensor, tokens_per_expert: torch.Tensor) -> torch.Tensor:
        """x_sorted [N,H] grouped by (local) expert; tokens_per_expert [E_local] (host list or tensor)."""
        if x_sorted.shape[0] > 0 and self._use_grouped(x_sorted):
            return self.forward_grouped(x_sorted, tokens_per_expert)
        counts = tokens_per_expert.tolist() if torch.is_tensor(tokens_per_expert) else list(tokens_per_expert)
        outs, off = [], 0
        for e, n in enumerate(counts):
            if n == 0:
                continue
            xe = x_sorted[off:off + n]
            h = self._act(xe @ self.mlp_op.gate_up_proj[e])
            outs.append(h @ self.mlp_op.down_proj[e])
            off += n
        if not outs:
            return x_sorted.new_zeros(0, self.H)
        return torch.cat(outs, dim=0)


class MoE(nn.Module):
    """router + dispatch + experts + combine.  Input/outputs ``[S_local, B, H]`` (SP) or ``[S, B, H]``."""

    def __init__(self, router: RouterBase, expert_mlps: ExpertMLPs, capacity_factor: float = 0.0,
                 normalize_top_k_affinities: bool = True, sequence_parallel_enabled: bool = False,
                 return_router_logits: bool = False, token_shuffle_group_size: int = 1):
        super().__init__()
        self.router, self.expert_mlps = router, expert_mlps
        self.capacity_factor = float(capacity_factor or 0.0)
        self.normalize = normalize_top_k_affinities
        self.sp = sequence_parallel_enabled and ps.get_tensor_model_parallel_size() > 1
        self.return_router_logits = return_router_logits
        self.token_shuffle_group_size = int(token_shuffle_group_size or 1)
        # created here (model construction happens on every rank at the same point): groups of consecutive DP ranks
        self._shuffle_group, self._shuffle_ranks = ps.get_token_shuffle_group(self.token_shuffle_group_size)

    # ---- expert-parallel exchange -------------------------------------------------------------------
    @staticmethod
    def _all_to_all(x: torch.Tensor, send_counts, recv_counts, group) -> torch.Tensor:
        out = x.new_empty((int(sum(recv_counts)),) + tuple(x.shape[1:]))
        if dist.get_backend(group) == "gloo":
            # gloo has no all-to-all: exchange through all_gather of padded buffers (CPU plumbing tests only)
            world = dist.get_world_size(group)
            r = dist.get_rank(group)
            mx = torch.tensor([x.shape[0]], dtype=torch.long)
            sizes = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
            dist.all_gather(sizes, mx, group=group)
            cap = int(max(s.item() for s in sizes))
            pad = x.new_zeros((cap,) + tuple(x.shape[1:]))
            pad[:x.shape[0]] = x
            bufs = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(bufs, pad, group=group)
            sc = torch.tensor(send_counts, dtype=torch.long)
            all_sc = [torch.zeros_like(sc) for _ in range(world)]
            dist.all_gather(all_sc, sc, group=group)
            off, pieces = 0, []
            for src in range(world):
                s_off = int(all_sc[src][:r].sum())
                n = int(all_sc[src][r])
                pieces.append(bufs[src][s_off:s_off + n])
            return torch.cat(pieces, 0) if pieces else out
        dist.all_to_all_single(out, x.contiguous(), output_split_sizes=list(map(int, recv_counts)),
                               input_split_sizes=list(map(int, send_counts)), group=group)
        return out

    def forward(self, hidden: torch.Tensor):
        tp_group = ps.get_tensor_model_parallel_group()
        tp = ps.group_size(tp_group)
        # Entering the tensor-parallel region: the experts are sharded on the FFN dimension, so every rank produces a
        # PARTIAL d(hidden); the entry mapping sums it in backward (reduce-scatter with SP, all-reduce without).
        if self.sp:
            hidden = mp.gather_from_sequence_parallel_region(hidden, True, tp_group)
        elif tp > 1:
            hidden = mp.copy_to_tensor_model_parallel_region(hidden, tp_group)
        shape = hidden.shape
        x = hidden.reshape(-1, shape[-1])
        T = x.shape[0]
        # token shuffling (reference X19 `token_shuffle_group_size`): tokens are randomly exchanged inside a group of DP
        # ranks before routing, so that a rank whose sequences favour a few experts does not overflow their capacity
        unshuffle = None
        G = len(self._shuffle_ranks)
        if self._shuffle_group is not None and G > 1 and T % G == 0:
            perm = torch.randperm(T, device=x.device)
            counts = [T // G] * G
            x = _A2A.apply(x.index_select(0, perm), counts, counts, self._shuffle_group)
            unshuffle = (perm, counts)
        E, k = self.router.num_experts, self.router.top_k
        # The router is replicated over the TP group (same tokens, same weights on every rank): its input gradient is
        # complete on each rank and must not be summed TP times by the entry mapping -> pre-scale it by 1/TP.
        logits, aff, idx = self.router(mp.scale_gradient(x, 1.0 / tp) if tp > 1 else x)
        w = aff.gather(-1, idx)                                   # [T,k]
        if self.normalize and k > 1:
            w = w / w.sum(-1, keepdim=True).clamp_min(1e-9)
        if tp > 1:
            # the affinities weight PARTIAL expert outputs (one FFN slice per rank): their gradient <d(out), y_partial> is
            # completed over the TP group, which also makes the router-weight gradient identical on every rank
            w = mp.copy_to_tensor_model_parallel_region(w, tp_group)
        flat_e = idx.reshape(-1)                                  # [T*k]
        flat_t = torch.arange(T, device=x.device).repeat_interleave(k)
        flat_w = w.reshape(-1)
        if self.capacity_factor > 0.0:
            cap = int(math.ceil(self.capacity_factor * T * k / E))
            # position of each assignment inside its expert queue (token order = priority)
            onehot = F.one_hot(flat_e, E)
            pos = (onehot.cumsum(0) - 1).gather(1, flat_e.unsqueeze(1)).squeeze(1)
            keep = pos < cap
            flat_e, flat_t, flat_w = flat_e[keep], flat_t[keep], flat_w[keep]
        order = torch.argsort(flat_e, stable=True)
        se, st, sw = flat_e[order], flat_t[order], flat_w[order]
        xs = x.index_select(0, st)
        counts = torch.bincount(se, minlength=E)
        ep = self.expert_mlps.ep
        if ep > 1:
            ep_group = ps.get_expert_model_parallel_group()
            El = self.expert_mlps.E_local
            send = counts.view(ep, El).sum(1).tolist()
            cnt_all = [torch.zeros_like(counts) for _ in range(ep)]
            dist.all_gather(cnt_all, counts, group=ep_group)
            r = self.expert_mlps.ep_rank
            recv_per_src = torch.stack([c[r * El:(r + 1) * El] for c in cnt_all])     # [ep, El]
            recv = recv_per_src.sum(1).tolist()
            xr = _A2A.apply(xs, send, recv, ep_group)
            # received layout: [src][local expert]; regroup by local expert
            src_sizes = recv_per_src.reshape(-1).tolist()
            chunks = list(torch.split(xr, src_sizes))
            regroup = [chunks[s * El + e] for e in range(El) for s in range(ep)]
            xe = torch.cat(regroup, 0) if regroup else xr
            ye = self.expert_mlps.forward_sorted(xe, recv_per_src.sum(0))
            # undo the regrouping
            sizes_e_major = [src_sizes[s * El + e] for e in range(El) for s in range(ep)]
            ychunks = list(torch.split(ye, sizes_e_major))
            back = [ychunks[e * ep + s] for s in range(ep) for e in range(El)]
            yr = torch.cat(back, 0) if back else ye
            ys = _A2A.apply(yr, recv, send, ep_group)
        else:
            ys = self.expert_mlps.forward_sorted(xs, counts)
        out = torch.zeros_like(x, dtype=torch.float32)
        out.index_add_(0, st, ys.float() * sw.unsqueeze(-1).float())
        out = out.to(x.dtype)
        if unshuffle is not None:
            perm, counts = unshuffle
            back = _A2A.apply(out, counts, counts, self._shuffle_group)       # tokens return to their home rank ...
            out = torch.empty_like(back).index_copy(0, perm, back)            # ... and to their original position
        out = out.view(shape)
        # experts are TP-sharded on the FFN dim => partial sums over TP
        if self.sp:
            out = mp.reduce_scatter_to_sequence_parallel_region(out, tp_group)
        else:
            out = mp.reduce_from_tensor_model_parallel_region(out, tp_group)
        if self.return_router_logits:
            return out, logits
        return out


class _A2A(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, send, recv, group):
        ctx.send, ctx.recv, ctx.group = send, recv, group
        return MoE._all_to_all(x, send, recv, group)

    @staticmethod
    def backward(ctx, g):
        return MoE._all_to_all(g.contiguous(), ctx.recv, ctx.send, ctx.group), None, None, None


class NeuronSwitchMLP(nn.Module):
    """MoE MLP for the Megatron GPT family (reference transformer.py:376-467): returns ``(output, None)`` like the
    dense MLP's ``(out, bias)`` and records the aux loss of the last forward."""

    def __init__(self, cfg, device=None):
        super().__init__()
        rk = dict(num_experts=cfg.num_moe_experts, top_k=cfg.moe_top_k, hidden_size=cfg.hidden_size,
                  act_fn=cfg.moe_router_activation, device=device, dtype=cfg.params_dtype)   # stored in the model dtype (the flat
        # optimizer buffers hold one dtype); the routing math itself runs in fp32 (RouterBase.logits)
        if cfg.moe_routing_algorithm == "sinkhorn":
            router = RouterSinkhorn(**rk, sinkhorn_iterations=cfg.moe_sinkhorn_iterations, sinkhorn_tol=cfg.moe_sinkhorn_tol)
        else:
            router = RouterTopK(**rk)
        act = "silu" if cfg.activation in ("swiglu", "fast-swiglu", "silu") else cfg.activation
        experts = ExpertMLPs(cfg.num_moe_experts, cfg.hidden_size, cfg.ffn_hidden_size, glu=cfg.moe_glu_mlp and cfg.is_glu,
                             activation=act, dtype=cfg.params_dtype, device=device, init_std=cfg.init_method_std,
                             use_cpu_initialization=cfg.use_cpu_initialization)
        self.moe = MoE(router, experts, capacity_factor=0.0 if cfg.moe_dropless else cfg.moe_capacity_factor,
                       normalize_top_k_affinities=cfg.moe_normalize_top_k_affinities,
                       sequence_parallel_enabled=cfg.sequence_parallel, return_router_logits=True,
                       token_shuffle_group_size=cfg.token_shuffle_group_size)
        self.cfg = cfg
        self.last_aux_loss = None

    def forward(self, hidden):
        out, logits = self.moe(hidden)
        self.last_aux_loss = load_balancing_loss_func(logits, self.cfg.num_moe_experts, self.cfg.moe_top_k) \
            if self.cfg.moe_aux_loss_coef > 0 else None
        return out, None
