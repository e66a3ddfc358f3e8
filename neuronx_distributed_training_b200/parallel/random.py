"""Tensor-parallel RNG tracker and seeding.

Capability parity: NxD ``parallel_layers.random`` (reference call sites
language_model.py:321-322, transformer.py:730-734,2528-2534) and the seeding rule
``seed + 100 * pp_rank`` (reference lightning_modules/model/megatron_init.py:63-75).

The default generator is seeded identically across a TP group (so replicated ops such as dropout
on non-sharded activations agree), while ``get_rng_tracker().fork()`` switches to a stream that
differs per TP rank (dropout on TP-sharded or sequence-sharded activations).
"""
from __future__ import annotations

import contextlib
import random as _py_random
from typing import Dict

import numpy as np
import torch

from . import state as ps

_MODEL_PARALLEL_RNG = "model-parallel-rng"


def _get_state(device_is_cuda: bool):
    return torch.cuda.get_rng_state() if device_is_cuda else torch.get_rng_state()


def _set_state(state, device_is_cuda: bool):
    if device_is_cuda:
        torch.cuda.set_rng_state(state)
    else:
        torch.set_rng_state(state)


class RNGStatesTracker:
    def __init__(self):
        self.states: Dict[str, torch.Tensor] = {}
        self.seeds = set()
        self.cuda = torch.cuda.is_available()

    def reset(self):
        self.states.clear()
        self.seeds.clear()

    def get_states(self):
        return {k: v.clone() for k, v in self.states.items()}

    def set_states(self, states):
        self.states = {k: v.clone() for k, v in states.items()}

    def add(self, name: str, seed: int):
        if seed in self.seeds:
            raise RuntimeError(f"seed {seed} already exists")
        if name in self.states:
            raise RuntimeError(f"rng state {name} already exists")
        self.seeds.add(seed)
        orig = _get_state(self.cuda)
        if self.cuda:
            torch.cuda.manual_seed(seed)
        else:
            torch.manual_seed(seed)
        self.states[name] = _get_state(self.cuda)
        _set_state(orig, self.cuda)

    @contextlib.contextmanager
    def fork(self, name: str = _MODEL_PARALLEL_RNG):
        if name not in self.states:
            # not seeded (e.g. unit tests without model_parallel_manual_seed): behave as identity
            yield
            return
        orig = _get_state(self.cuda)
        _set_state(self.states[name], self.cuda)
        try:
            yield
        finally:
            self.states[name] = _get_state(self.cuda)
            _set_state(orig, self.cuda)


_TRACKER = RNGStatesTracker()


def get_rng_tracker() -> RNGStatesTracker:
    return _TRACKER


def model_parallel_manual_seed(seed: int) -> None:
    """default stream: ``seed`` (same on all TP ranks); forked stream: ``seed + 2718 + tp_rank``."""
    _TRACKER.reset()
    _TRACKER.cuda = torch.cuda.is_available()
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    _TRACKER.add(_MODEL_PARALLEL_RNG, seed + 2718 + ps.get_tensor_model_parallel_rank())


def set_random_seed(seed: int) -> int:
    """Seed python / numpy / torch with ``seed + 100 * pp_rank`` and the TP tracker."""
    s = int(seed) + 100 * ps.get_pipeline_model_parallel_rank()
    _py_random.seed(s)
    np.random.seed(s % (2 ** 32))
    model_parallel_manual_seed(s)
    return s
