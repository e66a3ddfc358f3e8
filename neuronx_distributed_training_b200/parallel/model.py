"""``initialize_parallel_model``: build the (stage-local) model, wrap recompute regions, inject LoRA,
and expose ``run_train / run_eval / local_named_parameters``.

Capability parity: ``nxd.initialize_parallel_model`` + ``NxDPPModel`` (reference
hf_models/llama_model.py:94-103, lightning_modules/model/base.py:374-381, utils/model_utils.py:6-8).
"""
from __future__ import annotations

import logging
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence

import torch
import torch.nn as nn

from . import state as ps
from ..pipeline.partition import stage_layer_ranges
from ..pipeline.schedule import P2PCommunicator, PipelineSchedule
from ..utils.activation_checkpoint import apply_activation_checkpointing

log = logging.getLogger("nxdt")


class ParallelModel(nn.Module):
    """Wrapper around this rank's model chunk(s).

    PP == 1: ``self.chunks[0]`` is the whole model and ``forward`` delegates to it.
    PP  > 1: ``run_train(**global_batch)`` runs the 1F1B / interleaved schedule.
    State-dict keys are the *unwrapped, globally numbered* names (checkpoint layout requirement).
    """

    def __init__(self, chunks: Sequence[nn.Module], num_microbatches: int, micro_batch_size: int,
                 hidden_size: int, seq_len_fn: Optional[Callable[[Dict[str, Any]], int]] = None,
                 sequence_parallel: bool = False, dtype: torch.dtype = torch.bfloat16):
        super().__init__()
        self.chunks = nn.ModuleList(chunks)
        self.num_microbatches, self.micro_batch_size = num_microbatches, micro_batch_size
        self.hidden_size, self.sequence_parallel, self.act_dtype = hidden_size, sequence_parallel, dtype
        self.pp = ps.get_pipeline_model_parallel_size()
        self._sched: Optional[PipelineSchedule] = None
        self._sched_shape = None

    # ---- plain access -------------------------------------------------------------------------------
    @property
    def module(self) -> nn.Module:
        return self.chunks[0]

    def forward(self, *args, **kwargs):
        assert self.pp == 1, "call run_train/run_eval under pipeline parallelism"
        return self.chunks[0](*args, **kwargs)

    def local_named_parameters(self):
        seen = set()
        for c in self.chunks:
            for n, p in c.named_parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    yield n, p

    def local_parameters(self):
        return [p for _, p in self.local_named_parameters()]

    def state_dict(self, *args, **kwargs):
        out = {}
        for c in self.chunks:
            out.update(c.state_dict(*args, **kwargs))
        return out

    def load_state_dict(self, sd, strict: bool = True):
        missing, unexpected = [], set(sd.keys())
        for c in self.chunks:
            own = c.state_dict()
            sub = {k: v for k, v in sd.items() if k in own}
            r = c.load_state_dict(sub, strict=False)
            missing += list(r.missing_keys)
            unexpected -= set(sub.keys())
        if strict and (missing or unexpected):
            raise RuntimeError(f"missing {missing[:5]}, unexpected {list(unexpected)[:5]}")
        return torch.nn.modules.module._IncompatibleKeys(missing, list(unexpected))

    # ---- pipeline ---------------------------------------------------------------------------------------
    def _microbatches(self, batch: Dict[str, Any]) -> List[Dict[str, Any]]:
        mbs, n = self.micro_batch_size, self.num_microbatches
        return [{k: (v[i * mbs:(i + 1) * mbs] if torch.is_tensor(v) else v) for k, v in batch.items()} for i in range(n)]

    def _stage_forward(self, chunk_idx: int, mb: Dict[str, Any], hidden_in: Optional[torch.Tensor]):
        chunk = self.chunks[chunk_idx]
        seq_len = mb["input_ids"].shape[1]
        is_first = hidden_in is None
        kw = dict(seq_len=seq_len)
        if is_first:
            kw["input_ids"] = mb["input_ids"]
        else:
            kw["hidden"] = hidden_in
        if getattr(chunk, "lm_head", None) is not None:
            out = chunk(labels=mb["labels"], loss_mask=mb.get("loss_mask"), **kw)
            return out[0]
        return chunk(**kw)

    def _schedule(self, batch: Dict[str, Any]) -> PipelineSchedule:
        ids = batch["input_ids"]
        S = ids.shape[1]
        tp = ps.get_tensor_model_parallel_size()
        s_local = S // tp if self.sequence_parallel else S
        shape = (s_local, self.micro_batch_size, self.hidden_size)
        if self._sched is None or self._sched_shape != shape:
            dev = next(self.parameters()).device
            comm = P2PCommunicator(shape, self.act_dtype, dev)
            self._sched = PipelineSchedule(self._stage_forward, len(self.chunks), comm, self.num_microbatches)
            self._sched_shape = shape
        return self._sched

    def run_train(self, **batch) -> torch.Tensor:
        return self._schedule(batch).run(self._microbatches(batch), train=True)

    def run_eval(self, **batch) -> torch.Tensor:
        return self._schedule(batch).run(self._microbatches(batch), train=False)


def initialize_parallel_model(module, provider: Callable[..., nn.Module], transformer_layer_cls, num_layers: int,
                              activation_checkpoint=None, checkpoint_skip: Iterable[str] = (),
                              pipeline_cuts: Optional[Sequence[str]] = None, lora_config=None,
                              hidden_size: Optional[int] = None) -> ParallelModel:
    """``module`` is the task module (gives dims, precision, device, ``init_weights``).
    ``provider(device=..., layer_range=..., with_embedding=..., with_head=...)`` builds a chunk."""
    dims, pol = module.dims, module.precision
    dev = module.device
    pp, vpp, pp_rank = dims.pp, dims.vpp, ps.get_pipeline_model_parallel_rank()
    chunks: List[nn.Module] = []
    if pp == 1:
        chunks.append(provider(device=dev))
    else:
        if pipeline_cuts is not None and hasattr(pipeline_cuts, "__iter__"):
            pipeline_cuts = list(pipeline_cuts)
        ranges = stage_layer_ranges(num_layers, pp, vpp, pipeline_cuts)[pp_rank]
        for c, (lo, hi) in enumerate(ranges):
            first = pp_rank == 0 and c == 0
            last = pp_rank == pp - 1 and c == vpp - 1
            chunks.append(provider(device=dev, layer_range=(lo, hi), with_embedding=first, with_head=last))
    for ch in chunks:
        ch.pipeline_num_microbatches = module.num_microbatches
        if lora_config is not None:
            from ..models.lora import inject_lora
            inject_lora(ch, lora_config)
        wrapped = apply_activation_checkpointing(ch, activation_checkpoint, transformer_layer_cls, checkpoint_skip)
        if wrapped and ps.get_rank() == 0:
            log.info("activation checkpointing on %d modules (%s)", len(wrapped), activation_checkpoint)
    pm = ParallelModel(chunks, module.num_microbatches, module.micro_batch_size, hidden_size or 0,
                       sequence_parallel=module.sequence_parallel, dtype=pol.param_dtype)
    return pm
