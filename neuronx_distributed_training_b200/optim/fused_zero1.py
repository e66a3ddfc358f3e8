"""Fused ZeRO-1 step: gradient reduce-scatter (peer pulls) + partitioned AdamW + parameter all-gather (peer pushes)
as two sm_100a kernels over NVSwitch peer memory (``ops/csrc/zero1_fused.cu``), replacing
``reduce_scatter_tensor`` -> per-segment Adam launches -> ``all_gather_into_tensor``.

``enable_fused_zero1(optimizer)`` re-homes the optimizer's flat parameter / gradient buffers into symmetric memory
(peer-mapped over NVLink) and routes ``optimizer.step()`` through the fused path.  The only remaining collective is
the scalar grad-norm all-reduce.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .. import ops
from ..parallel import state as ps


class FusedZero1:
    def __init__(self, opt):
        import torch.distributed._symmetric_memory as symm_mem
        assert opt.zero1 and opt.dp_size > 1, "fused ZeRO-1 needs a data-parallel group of size > 1"
        assert opt.master is not None, "fused ZeRO-1 needs fp32 master weights (bf16 parameters)"
        assert opt.algorithm == "adamw" and opt.exp_avg.dtype == torch.float32
        assert opt.dp_size <= 8
        C = ops.C()
        assert C is not None and hasattr(C, "zero1_reduce"), "sm_100a extension with zero1 kernels required"
        self.group = opt.dp_group
        gname = self.group.group_name
        dev = opt.device
        pbytes = opt.param_buffer.numel() * opt.param_buffer.element_size()
        gbytes = opt.grad_buffer.numel() * opt.grad_buffer.element_size()
        self.hs_bytes = 4096
        total = self.hs_bytes + pbytes + gbytes
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=dev)
        self.buf[: self.hs_bytes].zero_()
        new_p = self.buf[self.hs_bytes: self.hs_bytes + pbytes].view(opt.param_buffer.dtype)
        new_g = self.buf[self.hs_bytes + pbytes:].view(opt.grad_buffer.dtype)
        new_p.copy_(opt.param_buffer)
        new_g.copy_(opt.grad_buffer)
        opt.rebind_buffers(new_p, new_g)
        self.hdl = symm_mem.rendezvous(self.buf, gname)
        base = [int(p) for p in self.hdl.buffer_ptrs]
        self.hs_ptrs = base
        self.param_ptrs = [b + self.hs_bytes for b in base]
        self.grad_ptrs = [b + self.hs_bytes + pbytes for b in base]
        self.gshard = torch.empty(opt.shard_size, dtype=torch.float32, device=dev)
        self.done_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
        self.epoch = 0
        # TP-replicated segments are counted once (on TP rank 0) in the global norm
        mask = torch.ones(opt.shard_size // 64, dtype=torch.uint8, device=dev)
        if ps.get_tensor_model_parallel_rank() != 0:
            for s, lo, hi in opt._seg_ranges():
                if s.tp_dup:
                    mask[lo // 64: (hi + 63) // 64] = 0
        self.count_mask = mask
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    @torch.no_grad()
    def step(self, opt):
        C = ops.C()
        self.epoch += 1
        opt._sync_sequence_parallel_grads()
        scale = 1.0 / opt.dp_size if opt.average_over_dp else 1.0
        ops._count(2)
        sq = C.zero1_reduce(self.grad_ptrs, self.param_ptrs, self.hs_ptrs, opt.dp_rank, self.epoch, self.gshard, opt.shard_lo,
                            opt.grad_buffer.dtype == torch.float32, scale * opt.extra_grad_scale, self.count_mask)
        dist.all_reduce(sq, group=self.group)
        for grp in (ps.get_tensor_model_parallel_group(), ps.get_pipeline_model_parallel_group()):
            if ps.group_size(grp) > 1:
                dist.all_reduce(sq, group=grp)
        norm = sq.sqrt()
        opt.grad_norm = norm
        clip = torch.clamp(opt.max_grad_norm / (norm + 1e-6), max=1.0).float().reshape(1) if opt.max_grad_norm is not None \
            else torch.empty(0, device=opt.device)
        segs = opt._seg_ranges()
        lo = [l for _, l, _ in segs]
        hi = [h for _, _, h in segs]
        lr = [float(opt.param_groups[s.group_idx]["lr"]) for s, _, _ in segs]
        wd = [float(opt.param_groups[s.group_idx]["weight_decay"]) for s, _, _ in segs]
        g0 = opt.param_groups[0]
        b1, b2 = g0["betas"]
        t = opt.step_count
        C.zero1_adam_push(self.grad_ptrs, self.param_ptrs, self.hs_ptrs, opt.dp_rank, self.epoch, opt.master, opt.exp_avg,
                          opt.exp_avg_sq, self.gshard, opt.shard_lo, lo, hi, lr, wd, float(b1), float(b2), float(g0["eps"]),
                          1.0 - b1 ** t, 1.0 - b2 ** t, clip, bool(opt.stochastic_rounding), t * 7919, self.done_ctr)


def enable_fused_zero1(optimizer) -> bool:
    """Returns True when the fused path was attached."""
    if not torch.cuda.is_available() or not getattr(optimizer, "zero1", False) or optimizer.dp_size <= 1:
        return False
    if optimizer.master is None or optimizer.algorithm != "adamw":
        return False
    optimizer._fused_zero1 = FusedZero1(optimizer)
    return True
