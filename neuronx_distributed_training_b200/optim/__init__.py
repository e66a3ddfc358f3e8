"""Optimizer / scheduler registries (names the YAMLs use; reference optim/__init__.py:11-12)."""
from .distributed import DistributedOptimizer
from .lr_schedulers import (LRScheduler, LinearAnnealingWithWarmUp, CosineAnnealing, WarmupAnnealing,
                            get_scheduler, register_scheduler, prepare_lr_scheduler)

# name -> (algorithm, extra kwargs).  All run through the flat DistributedOptimizer.
_OPTIMIZERS = {
    "adamw": ("adamw", {}),
    "adamw_fp32OptState": ("adamw", {"state_dtype_fp32": True}),
    "fused_adam": ("adamw", {}),
    "adam": ("adam", {}),
    "sgd": ("sgd", {}),
    "adamax": ("adamax", {}),
    "adagrad": ("adagrad", {}),
    "rmsprop": ("rmsprop", {}),
    "novograd": ("novograd", {}),
    "adafactor": ("adafactor", {}),
}


def register_optimizer(name, algorithm, **extra):
    """Make ``model.optim.name: <name>`` available.  ``algorithm`` is the name of a built-in update rule (``adamw``,
    ``adam``, ``sgd``, ``adamax``, ``adagrad``, ``rmsprop``, ``novograd``) or a callable
    ``fn(param32, grad32, state, group, step)`` that updates the flat fp32 parameter view in place; ``state`` holds two
    persistent buffers (``exp_avg``, ``exp_avg_sq``) and ``group`` the YAML hyper-parameters (``lr``, ``weight_decay``,
    ``betas``, ``eps``, ...).  Either way the update runs inside the flat ``DistributedOptimizer``, i.e. with ZeRO-1
    sharding, fp32 master weights / gradient accumulation, global-norm clipping and checkpointing for free."""
    _OPTIMIZERS[name] = (algorithm, extra)


def get_optimizer(name):
    if name not in _OPTIMIZERS:
        raise KeyError(f"unknown optimizer '{name}'; registered: {sorted(_OPTIMIZERS)}")
    return _OPTIMIZERS[name]
