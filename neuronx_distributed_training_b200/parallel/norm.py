"""Sequence-parallel aware norms.

Capability parity: NxD ``LayerNorm`` / ``RMSNorm`` (reference fused_layer_norm.py:8-15,
modeling_llama.py:145-161).  Under sequence parallelism every TP rank normalises different tokens, so the
weight gradient is only a partial sum: parameters are tagged ``sequence_parallel_enabled`` and the
optimizer wrapper all-reduces those grads over the TP group before the step.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class RMSNorm(nn.Module):
    def __init__(self, hidden_size: int, eps: float = 1e-5, sequence_parallel_enabled: bool = False,
                 dtype: torch.dtype = torch.float32, device=None):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size, dtype=dtype, device=device))
        self.variance_epsilon = eps
        self.weight.sequence_parallel_enabled = sequence_parallel_enabled

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.rmsnorm(x, self.weight, self.variance_epsilon)

    def extra_repr(self):
        return f"{tuple(self.weight.shape)}, eps={self.variance_epsilon}"


class LayerNorm(nn.Module):
    """Affine LayerNorm; ``one_plus`` gives the Megatron "layernorm1p" (weight stored as w-1)."""

    def __init__(self, hidden_size: int, eps: float = 1e-5, sequence_parallel_enabled: bool = False,
                 bias: bool = True, one_plus: bool = False, dtype: torch.dtype = torch.float32, device=None):
        super().__init__()
        init = torch.zeros if one_plus else torch.ones
        self.weight = nn.Parameter(init(hidden_size, dtype=dtype, device=device))
        self.bias = nn.Parameter(torch.zeros(hidden_size, dtype=dtype, device=device)) if bias else None
        self.eps, self.one_plus = eps, one_plus
        self.weight.sequence_parallel_enabled = sequence_parallel_enabled
        if self.bias is not None:
            self.bias.sequence_parallel_enabled = sequence_parallel_enabled

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.layernorm(x, self.weight, self.bias, self.eps, self.one_plus)


def get_norm(kind: str, hidden_size: int, eps: float, sequence_parallel_enabled: bool, dtype, device=None):
    kind = kind.lower()
    if kind == "rmsnorm":
        return RMSNorm(hidden_size, eps, sequence_parallel_enabled, dtype, device)
    if kind == "layernorm":
        return LayerNorm(hidden_size, eps, sequence_parallel_enabled, True, False, dtype, device)
    if kind == "layernorm1p":
        return LayerNorm(hidden_size, eps, sequence_parallel_enabled, True, True, dtype, device)
    raise ValueError(f"unknown normalization {kind}")
