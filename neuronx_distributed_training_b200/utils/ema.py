"""Exponential moving average of the parameters (the reference exposes NeMo's EMA through ``exp_manager.ema``,
reference utils/exp_manager.py:298-305).  The EMA lives on the optimizer's fp32 master shard, so under ZeRO-1 it is
sharded for free and costs one fused ``lerp`` per step."""
from __future__ import annotations

from typing import Optional

import torch


class EMA:
    def __init__(self, optimizer, decay: float = 0.999, every_n_steps: int = 1):
        self.opt, self.decay, self.every = optimizer, float(decay), int(every_n_steps)
        src = optimizer.master if optimizer.master is not None else optimizer.param_buffer[optimizer.shard_lo:optimizer.shard_hi]
        self.shadow = src.detach().float().clone()
        self._backup: Optional[torch.Tensor] = None

    def _src(self):
        o = self.opt
        return o.master if o.master is not None else o.param_buffer[o.shard_lo:o.shard_hi]

    @torch.no_grad()
    def update(self, step: int) -> None:
        if step % self.every == 0:
            self.shadow.lerp_(self._src().float(), 1.0 - self.decay)

    @torch.no_grad()
    def swap_in(self) -> None:
        """evaluate with the averaged weights (call ``swap_out`` afterwards)"""
        import torch.distributed as dist
        o = self.opt
        self._backup = o.param_buffer.clone()
        o.param_buffer[o.shard_lo:o.shard_hi].copy_(self.shadow)
        if o.zero1:
            dist.all_gather_into_tensor(o.param_buffer, o.param_buffer[o.shard_lo:o.shard_hi], group=o.dp_group)

    @torch.no_grad()
    def swap_out(self) -> None:
        if self._backup is not None:
            self.opt.param_buffer.copy_(self._backup)
            self._backup = None

    def state_dict(self):
        return {"shadow": self.shadow, "decay": self.decay}

    def load_state_dict(self, sd):
        self.shadow.copy_(sd["shadow"])
        self.decay = float(sd.get("decay", self.decay))
