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
}


def register_optimizer(name, algorithm, **extra):
    _OPTIMIZERS[name] = (algorithm, extra)


def get_optimizer(name):
    if name not in _OPTIMIZERS:
        raise KeyError(f"unknown optimizer '{name}'; registered: {sorted(_OPTIMIZERS)}")
    return _OPTIMIZERS[name]
