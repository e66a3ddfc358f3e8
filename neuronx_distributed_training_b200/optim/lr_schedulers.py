"""LR schedules, registered by the names the YAMLs use.

``LinearAnnealingWithWarmUp`` is the reference's own (reference optim/lr_schedulers.py:11-23: HF
linear-warmup / linear-decay); the rest are the NeMo schedule names listed in the reference docs
(docs/general/config_overview.rst:649-653), re-implemented from their published formulas.
All are stepped once per optimizer step.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional


class LRScheduler:
    """Minimal stateful scheduler: lr(step) = base_lr_of_group * factor(step) (or absolute via ``get_lr``)."""

    def __init__(self, optimizer, max_steps: Optional[int] = None, warmup_steps: int = 0, warmup_ratio: Optional[float] = None,
                 constant_steps: int = 0, constant_ratio: Optional[float] = None, min_lr: float = 0.0, **kwargs):
        self.optimizer = optimizer
        self.max_steps = int(max_steps) if max_steps is not None and max_steps > 0 else None
        if warmup_ratio is not None and self.max_steps:
            warmup_steps = int(warmup_ratio * self.max_steps)
        if constant_ratio is not None and self.max_steps:
            constant_steps = int(constant_ratio * self.max_steps)
        self.warmup_steps = int(warmup_steps or 0)
        self.constant_steps = int(constant_steps or 0)
        self.min_lr = float(min_lr or 0.0)
        self.extra = kwargs
        self.base_lrs: List[float] = [float(g.get("initial_lr", g["lr"])) for g in optimizer.param_groups]
        for g, b in zip(optimizer.param_groups, self.base_lrs):
            g.setdefault("initial_lr", b)
        self.last_step = 0
        self._apply()

    # subclasses implement this: learning rate for a group at the current step
    def lr_at(self, base_lr: float, step: int) -> float:
        raise NotImplementedError

    def _warm(self, base_lr: float, step: int) -> float:
        return base_lr * (step + 1) / (self.warmup_steps + 1)

    def get_lr(self) -> List[float]:
        return [self.lr_at(b, self.last_step) for b in self.base_lrs]

    def get_last_lr(self) -> List[float]:
        return [g["lr"] for g in self.optimizer.param_groups]

    def _apply(self):
        for g, lr in zip(self.optimizer.param_groups, self.get_lr()):
            g["lr"] = lr

    def step(self, step: Optional[int] = None):
        self.last_step = self.last_step + 1 if step is None else int(step)
        self._apply()

    def state_dict(self):
        return {"last_step": self.last_step, "base_lrs": list(self.base_lrs)}

    def load_state_dict(self, sd):
        self.last_step = int(sd["last_step"])
        if "base_lrs" in sd and len(sd["base_lrs"]) == len(self.base_lrs):
            self.base_lrs = list(sd["base_lrs"])
        self._apply()


class LinearAnnealingWithWarmUp(LRScheduler):
    """0 -> base over warmup_steps, then linearly to 0 at max_steps."""

    def lr_at(self, base_lr, step):
        if step < self.warmup_steps:
            return base_lr * step / max(1, self.warmup_steps)
        total = self.max_steps or step + 1
        return base_lr * max(0.0, (total - step) / max(1, total - self.warmup_steps))


class _WarmupHoldAnneal(LRScheduler):
    """warmup -> anneal over (max_steps - warmup - constant) -> hold min_lr for constant_steps."""

    def anneal(self, base_lr, t, T):
        raise NotImplementedError

    def lr_at(self, base_lr, step):
        if self.warmup_steps > 0 and step <= self.warmup_steps:
            return self._warm(base_lr, step)
        if self.max_steps is None:
            return base_lr
        decay_steps = self.max_steps - self.warmup_steps - self.constant_steps
        t = step - self.warmup_steps
        if t >= decay_steps or step > self.max_steps:
            return self.min_lr
        return self.anneal(base_lr, t, max(1, decay_steps))


class CosineAnnealing(_WarmupHoldAnneal):
    def anneal(self, base_lr, t, T):
        return self.min_lr + 0.5 * (base_lr - self.min_lr) * (1.0 + math.cos(math.pi * t / T))


class WarmupAnnealing(_WarmupHoldAnneal):
    def anneal(self, base_lr, t, T):
        return max(self.min_lr, base_lr * (1.0 - t / T))


class PolynomialDecayAnnealing(_WarmupHoldAnneal):
    def anneal(self, base_lr, t, T):
        power = float(self.extra.get("power", 1.0))
        return (base_lr - self.min_lr) * (1.0 - t / T) ** power + self.min_lr


class SquareAnnealing(_WarmupHoldAnneal):
    def anneal(self, base_lr, t, T):
        return (base_lr - self.min_lr) * (1.0 - t / T) ** 2 + self.min_lr


class SquareRootAnnealing(_WarmupHoldAnneal):
    def anneal(self, base_lr, t, T):
        return (base_lr - self.min_lr) * (1.0 - t / T) ** 0.5 + self.min_lr


class WarmupHoldPolicy(LRScheduler):
    def lr_at(self, base_lr, step):
        if self.warmup_steps > 0 and step <= self.warmup_steps:
            return self._warm(base_lr, step)
        return base_lr


class InverseSquareRootAnnealing(LRScheduler):
    def lr_at(self, base_lr, step):
        if self.warmup_steps > 0 and step <= self.warmup_steps:
            return self._warm(base_lr, step)
        return max(self.min_lr, base_lr * math.sqrt(max(1, self.warmup_steps) / max(1, step)))


class NoamAnnealing(LRScheduler):
    def lr_at(self, base_lr, step):
        d_model = float(self.extra.get("d_model", 1.0))
        w = max(1, self.warmup_steps)
        s = max(1, step)
        return max(self.min_lr, base_lr * d_model ** -0.5 * min(s ** -0.5, s * w ** -1.5))


class ConstantLR(LRScheduler):
    def lr_at(self, base_lr, step):
        return base_lr


_SCHEDULERS: Dict[str, Callable[..., LRScheduler]] = {}


def register_scheduler(name: str, cls) -> None:
    _SCHEDULERS[name] = cls


for _n, _c in [("LinearAnnealingWithWarmUp", LinearAnnealingWithWarmUp), ("CosineAnnealing", CosineAnnealing),
               ("WarmupAnnealing", WarmupAnnealing), ("PolynomialDecayAnnealing", PolynomialDecayAnnealing),
               ("PolynomialHoldDecayAnnealing", PolynomialDecayAnnealing), ("SquareAnnealing", SquareAnnealing),
               ("SquareRootAnnealing", SquareRootAnnealing), ("WarmupPolicy", WarmupHoldPolicy),
               ("WarmupHoldPolicy", WarmupHoldPolicy), ("InverseSquareRootAnnealing", InverseSquareRootAnnealing),
               ("NoamAnnealing", NoamAnnealing), ("constant", ConstantLR), ("ConstantLR", ConstantLR)]:
    register_scheduler(_n, _c)


def get_scheduler(name: str):
    if name not in _SCHEDULERS:
        raise KeyError(f"unknown scheduler '{name}'; registered: {sorted(_SCHEDULERS)}")
    return _SCHEDULERS[name]


def prepare_lr_scheduler(optimizer, sched_cfg, max_steps: Optional[int] = None) -> Optional[LRScheduler]:
    """``sched_cfg`` = the YAML ``model.optim.sched`` node (name + kwargs)."""
    if sched_cfg is None:
        return None
    d = sched_cfg.to_dict() if hasattr(sched_cfg, "to_dict") else dict(sched_cfg)
    name = d.pop("name")
    if d.get("max_steps") in (None, -1) and max_steps is not None:
        d["max_steps"] = max_steps
    return get_scheduler(name)(optimizer, **d)
