"""LoRA adapters for the tensor-parallel linears.

Capability parity: ``nxd.modules.lora.LoraConfig`` as configured by the reference (reference
hf_models/llama_model.py:51-65, conf/hf_llama3_8B_SFT_lora_config.yaml:109-124): rank / alpha / dropout / bias,
``target_modules`` in {qkv_proj, o_proj, gate_up_proj, down_proj, ...}, adapters sharded consistently with their parent
(column parent: A replicated, B output-sharded; row parent: A input-sharded, B replicated), base weights frozen,
adapter-only or merged checkpoints.

The low-rank path rides the parent's sequence-parallel layout: for a column parent the rank-r activation (not the
hidden state) is what gets all-gathered; for a row parent the rank-r partial sums are reduce-scattered.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..parallel import mappings as mp
from ..parallel import state as ps
from ..parallel.layers import (ColumnParallelLinear, RowParallelLinear, GQAQKVColumnParallelLinear,
                               set_tensor_model_parallel_attributes)


@dataclass
class LoraConfig:
    lora_rank: int = 16
    lora_alpha: float = 32.0
    lora_dropout: float = 0.05
    lora_bias: str = "none"                       # none | all | lora_only
    lora_verbose: bool = False
    target_modules: List[str] = field(default_factory=lambda: ["qkv_proj"])
    load_lora_from_ckpt: bool = False
    save_lora_base: bool = False
    merge_lora: bool = False
    save_lora_config_adapter: bool = True
    merge_sharded_lora: bool = False

    @staticmethod
    def from_config(node) -> "LoraConfig":
        d = node.to_dict() if hasattr(node, "to_dict") else dict(node)
        known = LoraConfig.__dataclass_fields__
        return LoraConfig(**{k: v for k, v in d.items() if k in known})

    @property
    def scaling(self) -> float:
        return float(self.lora_alpha) / float(self.lora_rank)

    def to_adapter_config(self, base_model_name_or_path: Optional[str] = None) -> Dict[str, object]:
        """The ``adapter_config.json`` that ``save_lora_config_adapter`` stores next to the adapter weights
        (HF-PEFT field names, so the adapter can be described to other tooling)."""
        return {"peft_type": "LORA", "task_type": "CAUSAL_LM", "r": int(self.lora_rank), "lora_alpha": float(self.lora_alpha),
                "lora_dropout": float(self.lora_dropout), "bias": str(self.lora_bias),
                "target_modules": list(self.target_modules), "base_model_name_or_path": base_model_name_or_path,
                "fan_in_fan_out": False, "inference_mode": False}


class LoraAdapter(nn.Module):
    """``delta(x) = scaling * B(A(dropout(x)))`` attached to a parallel linear."""

    def __init__(self, parent: nn.Module, cfg: LoraConfig):
        super().__init__()
        self.kind = "row" if isinstance(parent, RowParallelLinear) else "column"
        self.group = parent.group
        self.sp = bool(getattr(parent, "sequence_parallel_enabled", False))
        self.scaling = cfg.scaling
        self.dropout = nn.Dropout(cfg.lora_dropout) if cfg.lora_dropout > 0 else nn.Identity()
        w = _parent_weight(parent)
        out_local, in_local = w.shape
        r = cfg.lora_rank
        dt, dev = w.dtype, w.device
        self.lora_A = nn.Parameter(torch.empty(r, in_local, dtype=dt, device=dev))
        self.lora_B = nn.Parameter(torch.zeros(out_local, r, dtype=dt, device=dev))
        nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
        tp = ps.group_size(self.group)
        if self.kind == "column":
            # B's rows follow the parent's output layout (e.g. gate||up interleaved per rank: stride 2)
            set_tensor_model_parallel_attributes(self.lora_B, True, 0, int(getattr(w, "partition_stride", 1) or 1))
            self.lora_A.sequence_parallel_enabled = tp > 1        # replicated, partial grads => TP all-reduce
        else:
            set_tensor_model_parallel_attributes(self.lora_A, True, 1, 1)
            self.lora_B.sequence_parallel_enabled = tp > 1 and self.sp

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.dropout(x)
        if self.kind == "column":
            if self.sp:
                xa = F.linear(x, self.lora_A)
                xa = mp.gather_from_sequence_parallel_region(xa, True, self.group)
            else:
                xa = F.linear(mp.copy_to_tensor_model_parallel_region(x, self.group), self.lora_A)
            return F.linear(xa, self.lora_B) * self.scaling
        xa = F.linear(x, self.lora_A)
        if self.sp:
            xa = mp.reduce_scatter_to_sequence_parallel_region(xa, self.group)
        else:
            xa = mp.reduce_from_tensor_model_parallel_region(xa, self.group)
        return F.linear(xa, self.lora_B) * self.scaling

    def delta_weight(self) -> torch.Tensor:
        """this rank's shard of ``scaling * B @ A`` (same sharding as the parent weight)."""
        return (self.lora_B.float() @ self.lora_A.float()) * self.scaling


def _parent_weight(parent: nn.Module) -> torch.Tensor:
    if isinstance(parent, GQAQKVColumnParallelLinear):
        if not parent.fuse_qkv:
            raise ValueError("LoRA on qkv_proj requires model.fuse_qkv=True")
        return parent.weight_qkv
    return parent.weight


class LoraLinear(nn.Module):
    """Wraps a parallel linear: ``y = parent(x) + adapter(x)``; keeps the parent's parameter names under ``base_layer``
    out of the state dict by delegating (names stay ``<module>.weight`` + ``<module>.lora_A/lora_B``)."""

    def __init__(self, parent: nn.Module, cfg: LoraConfig):
        super().__init__()
        self.base_layer = parent
        self.adapter = LoraAdapter(parent, cfg)
        self.is_qkv = isinstance(parent, GQAQKVColumnParallelLinear)
        self.merged = False

    def forward(self, x):
        y = self.base_layer(x)
        if self.merged:
            return y
        d = self.adapter(x)
        if self.is_qkv:
            p = self.base_layer
            dq, dk, dv = torch.split(d, [p.q_local, p.kv_local, p.kv_local], dim=-1)
            return y[0] + dq, y[1] + dk, y[2] + dv
        if isinstance(y, tuple):        # skip_bias_add parents return (out, bias)
            return (y[0] + d,) + tuple(y[1:])
        return y + d

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("base_layer"), name)

    @torch.no_grad()
    def merge(self):
        if not self.merged:
            w = _parent_weight(self.base_layer)
            w.add_(self.adapter.delta_weight().to(w.dtype))
            self.merged = True

    @torch.no_grad()
    def unmerge(self):
        if self.merged:
            w = _parent_weight(self.base_layer)
            w.sub_(self.adapter.delta_weight().to(w.dtype))
            self.merged = False


def inject_lora(model: nn.Module, cfg: LoraConfig) -> List[str]:
    """Freeze the base model, wrap every module whose attribute name is in ``cfg.target_modules``."""
    targets = set(cfg.target_modules)
    for p in model.parameters():
        p.requires_grad_(False)
    wrapped = []
    for name, mod in list(model.named_modules()):
        for child_name, child in list(mod.named_children()):
            if child_name in targets and isinstance(child, (ColumnParallelLinear, RowParallelLinear, GQAQKVColumnParallelLinear)):
                setattr(mod, child_name, LoraLinear(child, cfg))
                wrapped.append(f"{name}.{child_name}" if name else child_name)
    if cfg.lora_bias in ("all", "lora_only"):
        for n, p in model.named_parameters():
            if n.endswith("bias") and (cfg.lora_bias == "all" or ".adapter." in n or "base_layer" in n):
                p.requires_grad_(True)
    if not wrapped:
        raise ValueError(f"LoRA target_modules {sorted(targets)} matched nothing")
    if cfg.lora_verbose and ps.get_rank() == 0:
        import logging
        log = logging.getLogger("nxdt")
        n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
        n_all = sum(p.numel() for p in model.parameters())
        log.info("LoRA: rank %d, alpha %g, dropout %g, bias %s; wrapped %d modules (%s%s); trainable parameters on this "
                 "rank %d of %d (%.3f %%)", cfg.lora_rank, cfg.lora_alpha, cfg.lora_dropout, cfg.lora_bias, len(wrapped),
                 ", ".join(wrapped[:4]), ", ..." if len(wrapped) > 4 else "", n_train, n_all, 100.0 * n_train / max(n_all, 1))
    return wrapped


def lora_state_dict(model: nn.Module) -> Dict[str, torch.Tensor]:
    return {k: v for k, v in model.state_dict().items() if ".adapter.lora_" in k}


def merge_lora(model: nn.Module) -> None:
    for m in model.modules():
        if isinstance(m, LoraLinear):
            m.merge()


def merged_state_dict(model: nn.Module) -> Dict[str, torch.Tensor]:
    """base-model key names (no ``base_layer`` / adapter entries) with the LoRA delta folded in."""
    merge_lora(model)
    try:
        out = {}
        for k, v in model.state_dict().items():
            if ".adapter." in k:
                continue
            out[k.replace(".base_layer.", ".")] = v.clone()
        return out
    finally:
        for m in model.modules():
            if isinstance(m, LoraLinear):
                m.unmerge()
