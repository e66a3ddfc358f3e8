"""Selective / full activation checkpointing by module class.

``"full"``  -> recompute every transformer layer; a list of class names -> recompute every instance of
those classes (reference llama_model.py:76-93, megatron_base_model.py:56-69); ``None`` -> off.
Wrapping replaces ``module.forward`` in place, so parameter names / state-dict keys are unchanged.
The TP RNG tracker state is captured and replayed so dropout under sequence parallelism recomputes
identically.
"""
from __future__ import annotations

import functools
from typing import Iterable, List, Optional, Sequence, Type, Union

import torch
import torch.nn as nn
from torch.utils.checkpoint import checkpoint as _torch_checkpoint

from ..parallel.random import get_rng_tracker


def _checkpointed_forward(orig_forward, *args, **kwargs):
    tracker = get_rng_tracker()
    saved = tracker.get_states()
    first = [True]

    def run(*a, **kw):
        if first[0]:
            first[0] = False
            return orig_forward(*a, **kw)
        # recomputation: replay the TP RNG streams from the state seen in the first pass
        cur = tracker.get_states()
        tracker.set_states(saved)
        try:
            return orig_forward(*a, **kw)
        finally:
            tracker.set_states(cur)

    if not torch.is_grad_enabled():
        return orig_forward(*args, **kwargs)
    return _torch_checkpoint(run, *args, use_reentrant=False, preserve_rng_state=True, **kwargs)


def checkpoint_wrap(module: nn.Module) -> nn.Module:
    if getattr(module, "_activation_checkpointed", False):
        return module
    orig = module.forward
    module.forward = functools.partial(_checkpointed_forward, orig)
    module._activation_checkpointed = True
    module._checkpoint_wrapped_forward = orig
    return module


def apply_activation_checkpointing(model: nn.Module, spec: Union[None, str, Sequence[Union[str, Type[nn.Module]]]],
                                   layer_cls: Optional[Type[nn.Module]] = None,
                                   skip: Iterable[str] = ()) -> List[str]:
    """Returns the qualified names of wrapped modules."""
    if spec is None:
        return []
    skip = set(skip)
    wrapped = []
    if isinstance(spec, str):
        if spec != "full":
            raise ValueError(f"activation checkpoint spec must be 'full', a class list or None; got {spec!r}")
        assert layer_cls is not None, "full checkpointing needs the transformer layer class"
        targets = {layer_cls.__name__}
    else:
        targets = {c if isinstance(c, str) else c.__name__ for c in spec}
    targets -= skip
    for name, mod in model.named_modules():
        if type(mod).__name__ in targets:
            # a module may implement its own (cheaper) recomputation, e.g. SwiGLU recomputed by its consumer
            own = getattr(mod, "enable_recompute", None)
            if own is None or not own():
                checkpoint_wrap(mod)
            wrapped.append(name)
    return wrapped
