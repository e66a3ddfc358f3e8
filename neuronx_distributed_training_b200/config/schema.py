"""Config validation + precision policy + derived batch arithmetic.

Behavioural parity with the reference orchestrator's ``process_config``
(reference: examples/training_orchestrator.py:25-137) and with the batch arithmetic of
``BaseModelModule.__init__`` (reference: src/.../lightning_modules/model/base.py:54-57), but
precision is expressed as an explicit policy object instead of environment variables.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Optional

import torch

from .loader import Config

# Keys that only make sense for the Neuron/XLA stack.  They are accepted (so reference YAMLs load
# unmodified) and ignored.
NEURON_ONLY_KEYS = (
    "compiler_flags", "compiler_cache_url", "aync_exec_max_inflight_requests",
    "neuron_rt_exec_timeout", "neuron_experimental_compress_rg",
)
NEURON_ONLY_MODEL_KEYS = ("transpose_nki_inputs", "use_cpu_initialization")
NEURON_ONLY_TRAINER_KEYS = ("lnc", "sequential_move_factor")


@dataclass
class PrecisionPolicy:
    """What dtype parameters live in, what the matmuls run in, and how the optimizer treats them."""
    type: str = "mixed_precision"
    param_dtype: torch.dtype = torch.bfloat16      # dtype of the model's nn.Parameters
    compute_dtype: torch.dtype = torch.bfloat16    # GEMM / attention dtype
    master_weights: bool = True                    # optimizer keeps fp32 master copies
    fp32_grad_acc: bool = True                     # gradients accumulate in fp32 "main_grad"
    stochastic_rounding: bool = False              # fp32 -> bf16 write-back uses SR
    autocast: bool = False                         # fp32 params + torch.autocast(bf16)
    reduce_dtype: torch.dtype = torch.bfloat16     # TP reduce-scatter accumulation dtype
    fp8_gemm: bool = False                         # block-scaled fp8 GEMM path (B200 extension)
    optimizer_state_dtype: torch.dtype = torch.float32

    @staticmethod
    def from_config(cfg: Config, zero1: bool = True) -> "PrecisionPolicy":
        p = cfg.get("precision", None) or Config({})
        t = str(p.get("type", "mixed_precision"))
        # the reference spells this two ways (orchestrator: mixed_precision_SR, module:
        # mixed_precisionSR); accept both.
        if t == "mixed_precision_SR":
            t = "mixed_precisionSR"
        pol = PrecisionPolicy(type=t)
        if t in ("mixed_precision", "mixed_precisionSR"):
            pol.master_weights = True
            pol.fp32_grad_acc = True
            pol.stochastic_rounding = t.endswith("SR")
        elif t == "bf16SR":
            pol.master_weights = False
            pol.fp32_grad_acc = False
            pol.stochastic_rounding = True
        elif t == "fp32":
            pol.param_dtype = torch.float32
            pol.compute_dtype = torch.float32
            pol.master_weights = False
            pol.fp32_grad_acc = False
            pol.reduce_dtype = torch.float32
        elif t == "autocast":
            pol.param_dtype = torch.float32
            pol.compute_dtype = torch.bfloat16
            pol.autocast = True
            pol.master_weights = False
            pol.fp32_grad_acc = False
        elif t == "manual":
            pol.master_weights = bool(p.get("master_weights", False))
            pol.fp32_grad_acc = bool(p.get("fp32_grad_acc", False))
            pol.stochastic_rounding = str(p.get("neuron_rt_stochastic_rounding_en", "0")) == "1"
            if str(p.get("xla_use_bf16", "0")) == "0" and str(p.get("xla_downcast_bf16", "0")) == "0" \
                    and not p.get("bf16_params", True):
                pol.param_dtype = torch.float32
        else:
            raise ValueError(f"unknown precision.type {t!r}")
        rd = p.get("parallel_layers_reduce_dtype", None)
        if rd is not None:
            pol.reduce_dtype = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16,
                                "fp32": torch.float32, "float32": torch.float32}[str(rd)]
        pol.fp8_gemm = bool(p.get("fp8_gemm", False))
        # The reference only honours master weights / fp32 grad acc together with ZeRO-1
        # (base.py:127-132).  We honour them regardless, ZeRO-1 merely shards them.
        osd = p.get("optimizer_state_dtype", None)
        if osd is not None:
            pol.optimizer_state_dtype = {"bf16": torch.bfloat16, "fp32": torch.float32}[str(osd)]
        return pol


@dataclass
class ParallelDims:
    world_size: int
    tp: int = 1
    pp: int = 1
    cp: int = 1
    ep: int = 1
    vpp: int = 1
    dp: int = field(init=False)

    def __post_init__(self):
        denom = self.tp * self.pp * self.cp
        if self.world_size % denom != 0:
            raise ValueError(f"world_size {self.world_size} not divisible by tp*pp*cp={denom}")
        self.dp = self.world_size // denom
        if self.ep > 1 and (self.dp % self.ep) != 0:
            raise ValueError(f"expert parallel size {self.ep} must divide dp={self.dp}")

    @staticmethod
    def from_config(cfg: Config, world_size: Optional[int] = None) -> "ParallelDims":
        ds = cfg.get("distributed_strategy", None) or Config({})
        if world_size is None:
            tr = cfg.get("trainer", Config({}))
            world_size = int(tr.get("devices", 1)) * int(tr.get("num_nodes", 1))
        vpp = ds.get("virtual_pipeline_model_parallel_size", 1) or 1
        return ParallelDims(
            world_size=world_size,
            tp=int(ds.get("tensor_model_parallel_size", 1) or 1),
            pp=int(ds.get("pipeline_model_parallel_size", 1) or 1),
            cp=int(ds.get("context_parallel_size", 1) or 1),
            ep=int(ds.get("expert_model_parallel_size", 1) or 1),
            vpp=int(vpp),
        )


def batch_arithmetic(cfg: Config, dims: ParallelDims):
    """returns (micro_batch_size, global_batch_size, num_microbatches, per_dp_batch)."""
    data = cfg.data
    mbs = int(data.micro_batch_size)
    gbs = int(data.global_batch_size)
    if gbs % (mbs * dims.dp) != 0:
        raise ValueError(f"global_batch_size {gbs} must be divisible by micro_batch_size*dp = {mbs}*{dims.dp}")
    return mbs, gbs, gbs // (mbs * dims.dp), gbs // dims.dp


def padded_vocab_size(vocab_size: int, divisible_by: int, tp: int) -> int:
    """reference: lightning_modules/data/base.py:66-89."""
    multiple = divisible_by * tp
    return ((vocab_size + multiple - 1) // multiple) * multiple


def process_config(cfg: Config, world_size: Optional[int] = None) -> Config:
    """Validate + normalise in place (returns cfg).  Mirrors the checks the reference performs in
    training_orchestrator.process_config and megatron_base_model validation."""
    ds = cfg.setdefault("distributed_strategy", {})
    if ds.get("virtual_pipeline_model_parallel_size", None) is None:
        ds["virtual_pipeline_model_parallel_size"] = 1

    # COMPILE=1 on Neuron is an AOT graph-extraction dry run.  There is nothing to compile ahead
    # of time here (eager CUDA + pre-built kernels); keep the observable behaviour: 10 steps, no
    # TensorBoard, no checkpoints.
    if os.environ.get("COMPILE", "0") == "1":
        cfg.trainer["max_steps"] = 10
        cfg.exp_manager["create_tensorboard_logger"] = False
        cfg.exp_manager["create_checkpoint_callback"] = False
    elif os.environ.get("TRAIN_ITERS") is not None:
        cfg.trainer["max_steps"] = int(os.environ["TRAIN_ITERS"])

    model = cfg.get("model", Config({}))
    moe = model.get("moe", None)
    if moe is not None:
        dropless = bool(moe.get("dropless", False))
        cf = float(moe.get("capacity_factor", 0.0) or 0.0)
        if dropless:
            act = str(model.get("activation", "swiglu")).lower()
            if act not in ("silu", "swiglu"):
                raise ValueError("dropless MoE supports only SiLU/SwiGLU activations")
            if not bool(moe.get("glu_mlp", True)):
                raise ValueError("dropless MoE requires glu_mlp=True")
            moe["capacity_factor"] = 0.0
        elif cf <= 0.0:
            raise ValueError("non-dropless MoE requires capacity_factor > 0")

    PrecisionPolicy.from_config(cfg)  # validates precision.type
    if world_size is not None:
        dims = ParallelDims.from_config(cfg, world_size)
        if "data" in cfg and "global_batch_size" in cfg.data:
            batch_arithmetic(cfg, dims)
        if dims.tp == 1 and bool(ds.get("sequence_parallel", False)):
            # reference forces SP off when TP == 1 (megatron_base_model.py:76-80)
            ds["sequence_parallel"] = False
    return cfg
