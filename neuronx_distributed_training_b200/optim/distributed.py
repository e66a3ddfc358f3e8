"""Flat-buffer mixed-precision optimizer with ZeRO-1 sharding, global grad-norm clipping and
sequence-parallel gradient sync.

Capability parity: NxD ``initialize_parallel_optimizer`` / ZeRO-1 wrapper and ``AdamW_FP32OptimParams``
(reference lightning_modules/model/base.py:303-308, optim/__init__.py:7-11).

B200-first design: all parameters of the local model shard live in ONE flat buffer per dtype and all
gradients in ONE flat (optionally fp32 ``main_grad``) buffer, ordered so that each (param-group,
TP-duplicate?) class is a contiguous segment.  Consequences:
  * the ZeRO-1 data-parallel reduce-scatter / all-gather are single collectives over the flat buffer
    (NVLS-friendly; or the fused sm_100a ``zero1_step`` kernel over NVSwitch peer memory);
  * the global grad-norm is <= 4 fused reductions, the AdamW update <= 4 fused launches regardless of the
    number of parameter tensors;
  * wgrad GEMMs accumulate directly into the fp32 main-grad views (``param.main_grad``).
The clip coefficient stays on the device (no host sync in the step).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .. import ops
from ..parallel import state as ps
from ..parallel.layers import param_is_tp_duplicate
from ..parallel.grad_hooks import ordered_main_grad_update


@dataclass
class _Seg:
    start: int
    end: int
    group_idx: int
    tp_dup: bool


def _align(n: int, a: int) -> int:
    return (n + a - 1) // a * a


class DistributedOptimizer:
    """AdamW / Adam / SGD / Adamax / Adagrad / RMSprop / NovoGrad over flat buffers.

    Args:
        param_groups: list of dicts ``{"params": [...], "weight_decay": .., "lr": ..}``.
        zero1: shard optimizer state over the data-parallel group.
        master_weights: keep fp32 master copies of (bf16) parameters.
        fp32_grad_acc: gradients accumulate in an fp32 ``main_grad`` buffer.
        max_grad_norm: global-norm clipping threshold (``trainer.gradient_clip_val``); None/0 = off.
    """

    def __init__(self, param_groups: Sequence[Dict[str, Any]], lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.01, algorithm: str = "adamw", momentum: float = 0.0,
                 zero1: bool = True, master_weights: bool = True, fp32_grad_acc: bool = True,
                 max_grad_norm: Optional[float] = 1.0, stochastic_rounding: bool = False,
                 state_dtype: torch.dtype = torch.float32, dp_group=None, grad_sync_groups: Optional[dict] = None,
                 average_over_dp: bool = True, extra_grad_scale: float = 1.0,
                 extra_defaults: Optional[Dict[str, Any]] = None, **unused):
        self.algorithm = algorithm if callable(algorithm) else algorithm.lower()
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, momentum=momentum)
        # algorithm-specific hyper-parameters from the YAML (rmsprop alpha, novograd grad_averaging, adafactor
        # clip_threshold / decay_rate / beta1 / scale_parameter, custom rules' own keys): visible as group[...]
        self.defaults.update(extra_defaults or {})
        self.param_groups: List[Dict[str, Any]] = []
        for g in param_groups:
            g = dict(g)
            g["params"] = list(g["params"])
            for k, v in self.defaults.items():
                g.setdefault(k, v)
            self.param_groups.append(g)
        if isinstance(dp_group, str) and dp_group == "self":
            self.dp_group = None            # explicitly no data-parallel reduction (e.g. experts with EP == DP)
        else:
            self.dp_group = dp_group if dp_group is not None else ps.get_data_context_parallel_group()
        self.dp_size = ps.group_size(self.dp_group)
        self.dp_rank = ps.group_rank(self.dp_group)
        self.zero1 = bool(zero1) and self.dp_size > 1
        self.zero1_requested = bool(zero1)
        self.master_weights = master_weights
        self.fp32_grad_acc = fp32_grad_acc
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm else None
        self.stochastic_rounding = stochastic_rounding
        self.state_dtype = state_dtype
        self.average_over_dp = average_over_dp
        self.extra_grad_scale = float(extra_grad_scale)   # e.g. 1/EP for expert-parallel parameters
        self.step_count = 0
        self.grad_norm: Optional[torch.Tensor] = None
        self._hooks = []
        self._fused_zero1 = None   # set by parallel.fused.enable_fused_zero1()
        self._build_flat()

    # ------------------------------------------------------------------------------------------
    def _build_flat(self):
        entries = []
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                if p.requires_grad:
                    entries.append((gi, param_is_tp_duplicate(p), p))
        if not entries:
            raise ValueError("optimizer got no trainable parameters")
        # contiguous segments per (group, tp_dup)
        entries.sort(key=lambda e: (e[0], e[1]))
        self.params: List[torch.nn.Parameter] = [e[2] for e in entries]
        self.param_group_idx: List[int] = [e[0] for e in entries]
        dev = self.params[0].device
        pdtype = self.params[0].dtype
        for p in self.params:
            if p.dtype != pdtype:
                raise ValueError("all trainable parameters must share one dtype (got %s and %s)" % (pdtype, p.dtype))
        self.device, self.param_dtype = dev, pdtype
        offs, cur = [], 0
        self.segments: List[_Seg] = []
        for gi, dup, p in entries:
            # 16-element alignment keeps every param view 32B-aligned for vector / TMA access
            cur = _align(cur, 64)
            if not self.segments or (self.segments[-1].group_idx, self.segments[-1].tp_dup) != (gi, dup):
                self.segments.append(_Seg(cur, cur, gi, dup))
            offs.append(cur)
            cur += p.numel()
            self.segments[-1].end = cur
        self.total = _align(cur, 64 * max(1, self.dp_size) * 4)
        self.offsets = offs
        self.param_buffer = torch.zeros(self.total, dtype=pdtype, device=dev)
        gdtype = torch.float32 if self.fp32_grad_acc else pdtype
        self.grad_buffer = torch.zeros(self.total, dtype=gdtype, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                view = self.param_buffer[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.main_grad = self.grad_buffer[o:o + p.numel()].view(p.shape)
                p.grad_added_to_main_grad = False
                p.grad = None
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook()))
        # shard
        if self.zero1:
            self.shard_size = self.total // self.dp_size
            self.shard_lo = self.dp_rank * self.shard_size
        else:
            self.shard_size, self.shard_lo = self.total, 0
        self.shard_hi = self.shard_lo + self.shard_size
        n = self.shard_size
        need_master = self.master_weights and pdtype != torch.float32
        self.master = self.param_buffer[self.shard_lo:self.shard_hi].float() if need_master else None
        self.exp_avg = torch.zeros(n, dtype=self.state_dtype, device=dev)
        # a callable algorithm is a user-registered update rule (optim.register_optimizer): it gets both state buffers
        two_buffers = callable(self.algorithm) or self.algorithm in ("adamw", "adam", "adamax", "adagrad", "rmsprop", "adafactor")
        self.exp_avg_sq = torch.zeros(n, dtype=self.state_dtype, device=dev) if two_buffers else None
        # NovoGrad keeps ONE second-moment scalar per tensor (layer-wise normalisation)
        self.layer_v = torch.zeros(len(self.params), dtype=torch.float32, device=dev) if self.algorithm == "novograd" else None
        self.af_rc = None
        if self.algorithm == "adafactor":
            self._adafactor_setup()
        self.state = {"step": 0}

    def rebind_buffers(self, new_param_buffer: torch.Tensor, new_grad_buffer: torch.Tensor) -> None:
        """Move the flat buffers to new storage (e.g. symmetric memory); contents must already be copied."""
        assert new_param_buffer.numel() == self.total and new_grad_buffer.numel() == self.total
        self.param_buffer, self.grad_buffer = new_param_buffer, new_grad_buffer
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                p.data = self.param_buffer[o:o + p.numel()].view(p.shape)
                p.main_grad = self.grad_buffer[o:o + p.numel()].view(p.shape)

    @staticmethod
    def _make_hook():
        def hook(p):
            if p.grad is not None:
                g = p.grad
                # ordered across streams when micro-batches run on several of them (parallel/grad_hooks.py)
                ordered_main_grad_update(g.device, lambda: p.main_grad.add_(g.to(p.main_grad.dtype)))
                p.grad = None
                cb = getattr(p, "_grad_ready_cb", None)
                if cb is not None:
                    cb(p)
        return hook

    # ------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = True):
        self.grad_buffer.zero_()
        for p in self.params:
            p.grad = None
            p.grad_added_to_main_grad = False

    def _sync_sequence_parallel_grads(self):
        """Sum grads of SP-tagged replicated params (norm weights, row-parallel biases) over TP."""
        tp_group = ps.get_tensor_model_parallel_group()
        if ps.group_size(tp_group) <= 1:
            return
        sp = [p for p in self.params if getattr(p, "sequence_parallel_enabled", False)]
        if not sp:
            return
        flat = torch.cat([p.main_grad.reshape(-1) for p in sp])
        dist.all_reduce(flat, group=tp_group)
        o = 0
        for p in sp:
            p.main_grad.copy_(flat[o:o + p.numel()].view(p.shape))
            o += p.numel()

    # ---- data-parallel gradient reduction overlapped with the last micro-batch's backward -----------------
    def enable_grad_reduce_overlap(self, bucket_mb: float = 256.0) -> bool:
        """Split the flat gradient buffer into buckets (cut at the ZeRO-1 shard boundaries, so every bucket has ONE
        owner) and reduce a bucket to its owner as soon as every parameter in it has its final gradient -- while
        backward is still running (`bucket_size_collectives`, reference training_orchestrator.py:42).  The shard layout,
        the optimizer state and the checkpoints are unchanged: the owner receives the sum in place with
        ``dist.reduce`` (``all_reduce`` without ZeRO-1).  Buckets that hold parameters whose gradient needs another
        reduction first (sequence-parallel norm weights: TP sum) or that are written more than once per backward (tied
        weights) are reduced at step time."""
        if self.dp_size <= 1 or self._fused_zero1 is not None:
            return False
        esize = self.grad_buffer.element_size()
        per = max(64, int(bucket_mb * (1 << 20) / esize) // 64 * 64)
        cuts = set(range(0, self.total, per)) | {self.total}
        if self.zero1:
            cuts |= {r * self.shard_size for r in range(self.dp_size + 1)}
        cuts = sorted(cuts)
        self._buckets = [(lo, hi, (lo // self.shard_size) if self.zero1 else None) for lo, hi in zip(cuts[:-1], cuts[1:]) if hi > lo]
        starts = [b[0] for b in self._buckets]
        import bisect
        self._bucket_params = [0] * len(self._buckets)
        self._bucket_late = [False] * len(self._buckets)
        for o, p_ in zip(self.offsets, self.params):
            first = bisect.bisect_right(starts, o) - 1
            last = bisect.bisect_right(starts, o + p_.numel() - 1) - 1
            p_._grad_buckets = list(range(first, last + 1))
            late = bool(getattr(p_, "sequence_parallel_enabled", False)) or bool(getattr(p_, "_multi_use", False)) \
                or bool(getattr(p_, "kv_replica_sync", False))
            for b in p_._grad_buckets:
                self._bucket_params[b] += 1
                self._bucket_late[b] = self._bucket_late[b] or late
            p_._grad_ready_cb = self._on_grad_ready
        self._overlap = True
        self._armed = False
        return True

    def start_grad_sync_window(self) -> None:
        """Call right before the backward of the LAST micro-batch of the step."""
        if not getattr(self, "_overlap", False):
            return
        self._armed = True
        self._pending = list(self._bucket_params)
        self._launched = [False] * len(self._buckets)
        self._handles = []

    def _on_grad_ready(self, p) -> None:
        if not getattr(self, "_armed", False):
            return
        for b in p._grad_buckets:
            self._pending[b] -= 1
            if self._pending[b] == 0 and not self._bucket_late[b]:
                self._launch_bucket(b)

    def _launch_bucket(self, b: int) -> None:
        lo, hi, owner = self._buckets[b]
        t = self.grad_buffer[lo:hi]
        scale = self.extra_grad_scale / (self.dp_size if self.average_over_dp else 1.0)
        if scale != 1.0:
            t.mul_(scale)
        if owner is None:
            h = dist.all_reduce(t, group=self.dp_group, async_op=True)
        else:
            dst = dist.get_global_rank(self.dp_group, owner) if self.dp_group is not None else owner
            h = dist.reduce(t, dst=dst, group=self.dp_group, async_op=True)
        self._handles.append(h)
        self._launched[b] = True

    def _finish_overlapped_reduce(self) -> torch.Tensor:
        for b in range(len(self._buckets)):
            if not self._launched[b]:
                self._launch_bucket(b)
        for h in self._handles:
            h.wait()
        self._armed = False
        self._handles = []
        return self.grad_buffer[self.shard_lo:self.shard_hi] if self.zero1 else self.grad_buffer

    def _reduce_over_dp(self) -> torch.Tensor:
        """Returns this rank's gradient shard (view) after the DP reduction."""
        if getattr(self, "_armed", False):
            return self._finish_overlapped_reduce()
        if self.extra_grad_scale != 1.0:
            self.grad_buffer.mul_(self.extra_grad_scale)
        if self.dp_size == 1:
            return self.grad_buffer
        if self.average_over_dp:
            self.grad_buffer.div_(self.dp_size)
        if self.zero1:
            shard = self.grad_buffer[self.shard_lo:self.shard_hi]
            dist.reduce_scatter_tensor(shard, self.grad_buffer, group=self.dp_group)
            return shard
        dist.all_reduce(self.grad_buffer, group=self.dp_group)
        return self.grad_buffer

    def _seg_ranges(self):
        """(seg, lo, hi) in *shard-local* coordinates for segments intersecting this rank's shard."""
        out = []
        for s in self.segments:
            lo, hi = max(s.start, self.shard_lo), min(s.end, self.shard_hi)
            if lo < hi:
                out.append((s, lo - self.shard_lo, hi - self.shard_lo))
        return out

    def _global_grad_norm(self, gshard: torch.Tensor) -> torch.Tensor:
        tp_rank = ps.get_tensor_model_parallel_rank()
        total = torch.zeros((), dtype=torch.float32, device=self.device)
        for s, lo, hi in self._seg_ranges():
            if s.tp_dup and tp_rank != 0:
                continue
            total = total + ops.sq_norm(gshard[lo:hi])
        # KV-head replicas (TP larger than the number of KV heads): elements another rank already counts
        for o, p in zip(self.offsets, self.params):
            dup = getattr(p, "norm_duplicate_range", None)
            if dup is not None:
                lo, hi = max(o + dup[0], self.shard_lo), min(o + dup[1], self.shard_hi)
                if lo < hi:
                    total = total - ops.sq_norm(gshard[lo - self.shard_lo:hi - self.shard_lo].contiguous())
        if dist.is_initialized():
            if self.zero1:
                dist.all_reduce(total, group=self.dp_group)
            for grp in (ps.get_tensor_model_parallel_group(), ps.get_pipeline_model_parallel_group()):
                if ps.group_size(grp) > 1:
                    dist.all_reduce(total, group=grp)
        return total.sqrt()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if self._fused_zero1 is not None:
            self.step_count += 1
            self.state["step"] = self.step_count
            self._fused_zero1.step(self)
            return loss
        sq = self.prepare_step()
        self.apply_step(sq.sqrt())
        return loss

    def prepare_step(self, extra_norm_groups=()) -> torch.Tensor:
        """Phase 1: SP grad sync + DP reduction; returns the squared global grad norm of THIS optimizer's params."""
        self.step_count += 1
        self.state["step"] = self.step_count
        self._sync_sequence_parallel_grads()
        self._gshard = self._reduce_over_dp()
        sq = self._global_grad_norm(self._gshard) ** 2
        for grp in extra_norm_groups:
            if ps.group_size(grp) > 1:
                dist.all_reduce(sq, group=grp)
        return sq

    def apply_step(self, norm: torch.Tensor) -> None:
        """Phase 2: clip with the (possibly combined) global norm, update the shard, all-gather params."""
        self.grad_norm = norm
        clip = torch.clamp(self.max_grad_norm / (norm + 1e-6), max=1.0) if self.max_grad_norm is not None else None
        gshard = self._gshard
        pshard = self.param_buffer[self.shard_lo:self.shard_hi]
        if self.algorithm == "novograd":
            self._novograd_layer_moments(gshard, clip)
        if self.algorithm == "adafactor":
            self._adafactor_step(gshard, pshard, clip)
        else:
            for s, lo, hi in self._seg_ranges():
                g = self.param_groups[s.group_idx]
                self._update(pshard[lo:hi], gshard[lo:hi], lo, hi, g, clip)
        if self.zero1:
            dist.all_gather_into_tensor(self.param_buffer, pshard, group=self.dp_group)
        self._gshard = None

    def _update(self, p, g, lo, hi, group, clip):
        lr, wd = float(group["lr"]), float(group["weight_decay"])
        master = self.master[lo:hi] if self.master is not None else None
        if callable(self.algorithm):
            # user-registered element-wise update rule on the flat fp32 views of this rank's shard:
            #   fn(param32, grad32, state, group, step) updates param32 in place; state["exp_avg"] / state["exp_avg_sq"]
            #   are two persistent buffers of the same shape (checkpointed and ZeRO-1-sharded like Adam's moments)
            p32 = master if master is not None else p.float()
            gg = g.float() * (clip if clip is not None else 1.0)
            self.algorithm(p32, gg, {"exp_avg": self.exp_avg[lo:hi], "exp_avg_sq": self.exp_avg_sq[lo:hi]}, group,
                           self.step_count)
            if master is not None or p32.data_ptr() != p.data_ptr():
                p.copy_(p32)
        elif self.algorithm in ("adamw", "adam"):
            b1, b2 = group["betas"]
            scale = clip
            if self.algorithm == "adam" and wd != 0.0:
                # classic L2: fold decay into the gradient
                g = g.float() + wd * (master if master is not None else p.float())
                wd_eff = 0.0
            else:
                wd_eff = wd
            ops.adamw_step_flat(master, p, g, self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi], lr, b1, b2,
                                float(group["eps"]), wd_eff, self.step_count,
                                grad_scale=scale if scale is not None else 1.0,
                                stochastic_rounding=self.stochastic_rounding, seed=self.step_count * 7919 + lo)
        elif self.algorithm in ("adamax", "adagrad", "rmsprop"):
            gg = g.float() * (clip if clip is not None else 1.0)
            p32 = master if master is not None else p.float()
            if wd != 0.0:
                gg = gg + wd * p32
            eps = float(group["eps"])
            if self.algorithm == "adamax":
                b1, b2 = group["betas"]
                m, u = self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi]
                m.mul_(b1).add_(gg, alpha=1.0 - b1)
                torch.maximum(u.mul_(b2), gg.abs().add_(eps), out=u)
                p32.addcdiv_(m, u, value=-lr / (1.0 - b1 ** self.step_count))
            elif self.algorithm == "adagrad":
                acc = self.exp_avg_sq[lo:hi]
                acc.addcmul_(gg, gg)
                p32.addcdiv_(gg, acc.sqrt().add_(eps), value=-lr)
            else:   # rmsprop (alpha = betas[1]-style smoothing constant; optional momentum)
                alpha = float(group.get("alpha", 0.99))
                sq = self.exp_avg_sq[lo:hi]
                sq.mul_(alpha).addcmul_(gg, gg, value=1.0 - alpha)
                upd = gg / sq.sqrt().add_(eps)
                mom = float(group.get("momentum", 0.0))
                if mom > 0.0:
                    buf = self.exp_avg[lo:hi]
                    buf.mul_(mom).add_(upd)
                    upd = buf
                p32.add_(upd, alpha=-lr)
            if master is not None:
                p.copy_(p32)
            elif p32.data_ptr() != p.data_ptr():
                p.copy_(p32)
        elif self.algorithm == "novograd":
            b1, _ = group["betas"]
            p32 = master if master is not None else p.float()
            gg = g.float() * (clip if clip is not None else 1.0)
            denom = self._novograd_denom[lo:hi]
            upd = gg / denom
            if wd != 0.0:
                upd = upd + wd * p32
            if bool(group.get("grad_averaging", False)):
                upd = upd * (1.0 - b1)
            m = self.exp_avg[lo:hi]
            m.mul_(b1).add_(upd)
            p32.add_(m, alpha=-lr)
            p.copy_(p32)
        elif self.algorithm == "sgd":
            gg = g.float() * (clip if clip is not None else 1.0)
            p32 = master if master is not None else p.float()
            if wd != 0.0:
                gg = gg + wd * p32
            mom = float(group.get("momentum", 0.0))
            if mom > 0.0:
                buf = self.exp_avg[lo:hi]
                buf.mul_(mom).add_(gg)
                gg = buf
            p32.add_(gg, alpha=-lr)
            p.copy_(p32)
        else:
            raise ValueError(f"unsupported flat optimizer algorithm {self.algorithm}")

    def _novograd_layer_moments(self, gshard: torch.Tensor, clip) -> None:
        """Per-tensor ||g||^2 (summed over the ZeRO-1 shards that hold pieces of the tensor), EMA into ``layer_v`` and
        expansion to a per-element denominator for this rank's shard."""
        n = len(self.params)
        sq = torch.zeros(n, dtype=torch.float32, device=self.device)
        pieces = []
        for i, (o, p) in enumerate(zip(self.offsets, self.params)):
            lo, hi = max(o, self.shard_lo), min(o + p.numel(), self.shard_hi)
            if lo < hi:
                pieces.append((i, lo - self.shard_lo, hi - self.shard_lo))
                sq[i] = gshard[lo - self.shard_lo:hi - self.shard_lo].float().pow(2).sum()
        if clip is not None:
            sq = sq * clip * clip
        if self.zero1:
            dist.all_reduce(sq, group=self.dp_group)
        b2 = float(self.param_groups[0]["betas"][1])
        if self.step_count <= 1:
            self.layer_v.copy_(sq)
        else:
            self.layer_v.mul_(b2).add_(sq, alpha=1.0 - b2)
        eps = float(self.param_groups[0]["eps"])
        denom = torch.ones(self.shard_size, dtype=torch.float32, device=self.device)
        d = self.layer_v.sqrt() + eps
        for i, lo, hi in pieces:
            denom[lo:hi] = d[i]
        self._novograd_denom = denom

    # ---- Adafactor (Shazeer & Stern 2018; hyper-parameter names of the NeMo / HF implementations) -------------
    # Second moments of a >= 2-D tensor [..., r, c] are kept FACTORED: a row vector R [..., r] and a column vector
    # C [..., c].  Under ZeRO-1 a rank owns an arbitrary contiguous piece of the flattened tensor, so every rank adds the
    # row / column sums of its piece into one small vector that is all-reduced over the data-parallel group; R and C end
    # up replicated (they are tiny), everything element-wise stays sharded.  1-D tensors use the ordinary second moment
    # (``exp_avg_sq``); the optional first moment (``beta1``) lives in ``exp_avg``.  Like every TP implementation, the
    # statistics of a tensor-parallel weight are those of the local shard.
    def _adafactor_setup(self) -> None:
        self._af_meta = []          # per param: (factored, batch, rows, cols, offset of R in af_rc, offset of C)
        off = 0
        for p in self.params:
            shp = tuple(p.shape)
            if len(shp) >= 2:
                r, c = int(shp[-2]), int(shp[-1])
                b = p.numel() // (r * c)
                self._af_meta.append((True, b, r, c, off, off + b * r))
                off += b * (r + c)
            else:
                self._af_meta.append((False, 1, 1, int(p.numel()), -1, -1))
        self.af_rc = torch.zeros(max(off, 1), dtype=torch.float32, device=self.device)

    def _af_pieces(self):
        """(param index, lo, hi in shard-local coordinates, offset of the piece inside its tensor)"""
        out = []
        for i, (o, p) in enumerate(zip(self.offsets, self.params)):
            lo, hi = max(o, self.shard_lo), min(o + p.numel(), self.shard_hi)
            if lo < hi:
                out.append((i, lo - self.shard_lo, hi - self.shard_lo, lo - o))
        return out

    def _adafactor_step(self, gshard: torch.Tensor, pshard: torch.Tensor, clip) -> None:
        CH = 1 << 22                                             # elements per index_add chunk (bounds temporaries)
        t = self.step_count
        hp = self.param_groups[0]
        decay_rate = float(hp.get("decay_rate", -0.8))
        beta2t = 1.0 - math.pow(t, decay_rate)
        eps = hp.get("eps", None)
        eps1, eps2 = (float(eps[0]), float(eps[1])) if isinstance(eps, (tuple, list)) else (1e-30, 1e-3)
        clip_threshold = float(hp.get("clip_threshold", 1.0))
        beta1 = hp.get("beta1", None)
        scale_parameter = bool(hp.get("scale_parameter", True))
        pieces = self._af_pieces()
        g32 = gshard.float() if gshard.dtype != torch.float32 else gshard.clone()
        if clip is not None:
            g32.mul_(clip)
        # 1. row / column sums of g^2 + eps1 for the factored tensors
        part = torch.zeros_like(self.af_rc)
        for i, lo, hi, toff in pieces:
            fact, b, r, c, ro, co = self._af_meta[i]
            if not fact:
                continue
            for a in range(lo, hi, CH):
                e = min(hi, a + CH)
                idx = torch.arange(toff + (a - lo), toff + (e - lo), device=self.device)
                upd = g32[a:e].pow(2).add_(eps1)
                part.index_add_(0, ro + idx // c, upd)                               # row id = b * r + i
                part.index_add_(0, co + (idx // (r * c)) * c + idx % c, upd)         # column id = b * c + j
        if self.zero1:
            dist.all_reduce(part, group=self.dp_group)
        for i, (fact, b, r, c, ro, co) in enumerate(self._af_meta):
            if fact:
                part[ro:ro + b * r].div_(c)          # mean over the columns
                part[co:co + b * c].div_(r)          # mean over the rows
        self.af_rc.mul_(beta2t).add_(part, alpha=1.0 - beta2t)
        # 2. the unclipped update u (in place of g32) and the per-tensor sums of u^2 and p^2
        p32 = self.master if self.master is not None else pshard.float()
        stats = torch.zeros(2 * len(self.params), dtype=torch.float32, device=self.device)
        for i, lo, hi, toff in pieces:
            fact, b, r, c, ro, co = self._af_meta[i]
            if fact:
                R = self.af_rc[ro:ro + b * r].view(b, r)
                rfac = (R / R.mean(dim=1, keepdim=True)).rsqrt_().reshape(-1)
                cfac = self.af_rc[co:co + b * c].rsqrt()
                for a in range(lo, hi, CH):
                    e = min(hi, a + CH)
                    idx = torch.arange(toff + (a - lo), toff + (e - lo), device=self.device)
                    g32[a:e].mul_(rfac[idx // c]).mul_(cfac[(idx // (r * c)) * c + idx % c])
            else:
                v = self.exp_avg_sq[lo:hi]
                v.mul_(beta2t).add_(g32[lo:hi].pow(2).add_(eps1), alpha=1.0 - beta2t)
                g32[lo:hi].mul_(v.float().rsqrt())
            stats[2 * i] = g32[lo:hi].pow(2).sum()
            stats[2 * i + 1] = p32[lo:hi].float().pow(2).sum()
        if self.zero1:
            dist.all_reduce(stats, group=self.dp_group)
        # 3. RMS clipping of the update, parameter-scaled learning rate, optional momentum, decoupled weight decay
        for i, lo, hi, toff in pieces:
            grp = self.param_groups[self.param_group_idx[i]]
            n = float(self.params[i].numel())
            rms_u = (stats[2 * i] / n).sqrt()
            lr_t = float(grp["lr"])
            if scale_parameter:
                lr_t = lr_t * torch.clamp((stats[2 * i + 1] / n).sqrt(), min=eps2)
            u = g32[lo:hi]
            u.div_(torch.clamp(rms_u / clip_threshold, min=1.0)).mul_(lr_t)
            if beta1 is not None:
                m = self.exp_avg[lo:hi]
                m.mul_(float(beta1)).add_(u, alpha=1.0 - float(beta1))
                u = m.float()
            wd = float(grp["weight_decay"])
            if wd != 0.0:
                p32[lo:hi].mul_(1.0 - wd * lr_t)
            p32[lo:hi].sub_(u)
        if self.master is not None or p32.data_ptr() != pshard.data_ptr():
            pshard.copy_(p32)

    # ------------------------------------------------------------------------------------------
    def state_dict(self) -> Dict[str, Any]:
        """This rank's shard (every DP rank saves its own under ZeRO-1)."""
        sd = {
            "step": self.step_count, "algorithm": self.algorithm if isinstance(self.algorithm, str) else getattr(self.algorithm, "__name__", "custom"), "zero1": self.zero1,
            "dp_size": self.dp_size if self.zero1 else 1, "dp_rank": self.dp_rank if self.zero1 else 0,
            "total": self.total, "shard_lo": self.shard_lo, "shard_hi": self.shard_hi,
            "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "master": self.master,
            "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
            "layout": [(o, int(p.numel())) for o, p in zip(self.offsets, self.params)],
        }
        if self.layer_v is not None:
            sd["layer_v"] = self.layer_v
        if self.af_rc is not None:
            sd["adafactor_rc"] = self.af_rc          # replicated over the data-parallel group
        return sd

    def load_state_dict(self, sd: Dict[str, Any]) -> None:
        self.step_count = int(sd["step"])
        self.state["step"] = self.step_count
        if int(sd["total"]) != self.total or int(sd["shard_lo"]) != self.shard_lo or int(sd["shard_hi"]) != self.shard_hi:
            raise ValueError("optimizer shard layout mismatch: checkpoint was written with a different "
                             "parallel layout; use checkpoint.reshard.load_optimizer_state_any_dp() (DP change) or the converter")
        self.exp_avg.copy_(sd["exp_avg"])
        if self.exp_avg_sq is not None and sd.get("exp_avg_sq") is not None:
            self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        if self.layer_v is not None and sd.get("layer_v") is not None:
            self.layer_v.copy_(sd["layer_v"])
        if self.af_rc is not None and sd.get("adafactor_rc") is not None:
            self.af_rc.copy_(sd["adafactor_rc"])
        if self.master is not None:
            if sd.get("master") is not None:
                self.master.copy_(sd["master"])
            else:
                self.master.copy_(self.param_buffer[self.shard_lo:self.shard_hi].float())
        for g, sg in zip(self.param_groups, sd.get("param_groups", [])):
            for k, v in sg.items():
                g[k] = v

    def sync_master_from_params(self):
        """After loading model weights (``weight_init_only``) refresh the fp32 masters."""
        if self.master is not None:
            self.master.copy_(self.param_buffer[self.shard_lo:self.shard_hi].float())

    def master_params_state(self) -> Optional[torch.Tensor]:
        return self.master

    def memory_bytes(self) -> Dict[str, int]:
        def nb(t):
            return 0 if t is None else t.numel() * t.element_size()
        return {"params": nb(self.param_buffer), "grads": nb(self.grad_buffer), "master": nb(self.master),
                "exp_avg": nb(self.exp_avg), "exp_avg_sq": nb(self.exp_avg_sq)}


class ChainedOptimizer:
    """Several flat optimizers stepped as one (dense params reduced over the DP group, expert-parallel params over
    the expert-data-parallel group) with ONE global grad-norm / clip coefficient."""

    def __init__(self, optimizers, extra_norm_groups=None):
        self.optimizers = list(optimizers)
        self.extra_norm_groups = extra_norm_groups or [() for _ in self.optimizers]
        self.grad_norm = None

    @property
    def param_groups(self):
        return [g for o in self.optimizers for g in o.param_groups]

    @property
    def zero1(self):
        return any(o.zero1 for o in self.optimizers)

    @property
    def master(self):
        return self.optimizers[0].master

    def zero_grad(self, set_to_none: bool = True):
        for o in self.optimizers:
            o.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        sq = None
        for o, extra in zip(self.optimizers, self.extra_norm_groups):
            s = o.prepare_step(extra)
            sq = s if sq is None else sq + s
        norm = sq.sqrt()
        for o in self.optimizers:
            o.apply_step(norm)
        self.grad_norm = norm
        return loss

    def state_dict(self):
        return {"chained": [o.state_dict() for o in self.optimizers]}

    def load_state_dict(self, sd):
        for o, s in zip(self.optimizers, sd["chained"]):
            o.load_state_dict(s)

    def sync_master_from_params(self):
        for o in self.optimizers:
            o.sync_master_from_params()

    def memory_bytes(self):
        out = {}
        for o in self.optimizers:
            for k, v in o.memory_bytes().items():
                out[k] = out.get(k, 0) + v
        return out
