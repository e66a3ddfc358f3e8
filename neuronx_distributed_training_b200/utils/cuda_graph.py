"""CUDA-graph capture of one micro-batch forward + backward.

With tensor parallelism the per-rank kernels of an 8B model are tens of microseconds long, so an eager step is bound
by the Python / launch path rather than by the GPU.  ``MicrobatchGraph`` records the whole micro-batch (forward, loss,
backward with gradient accumulation into the flat ``main_grad`` buffer, NCCL collectives and the fused
compute+collective kernels) once and replays it for every following micro-batch; inputs are copied into static
buffers.  The reference gets the same effect from its ahead-of-time compiled XLA graphs (``neuron_parallel_compile``,
SURVEY X27); here it is opt-in: ``model.fusions.cuda_graph: True`` (or ``NXDT_CUDA_GRAPH=1``).

Requirements (checked where possible): static shapes, no host synchronisation inside the model, no
checkpoint-wrapped modules that snapshot the CUDA RNG state.  Status (measured): correct on one GPU and with NCCL
tensor parallelism on two; NOT yet with the fused TP kernels (see ``capture``), although those already keep their
epoch on the device and offer ``FusedTP.graph_boundary()`` for it.
"""
from __future__ import annotations

import logging
import os
from typing import Callable, Dict, Optional

import torch

log = logging.getLogger("nxdt")


def _fused_tp_objects():
    try:
        from ..parallel import fused
        return list(fused._FTP_CACHE.values())
    except Exception:  # noqa
        return []


class MicrobatchGraph:
    """``fn(static_batch) -> loss`` must run forward AND backward and return the (already scaled) loss tensor."""

    def __init__(self, fn: Callable[[Dict[str, torch.Tensor]], torch.Tensor]):
        self.fn = fn
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static: Dict[str, torch.Tensor] = {}
        self.meta: Dict[str, object] = {}
        self.loss: Optional[torch.Tensor] = None
        self.native_launches = 0

    def matches(self, mb: Dict[str, object]) -> bool:
        for k, v in mb.items():
            if torch.is_tensor(v):
                s = self.static.get(k)
                if s is None or s.shape != v.shape or s.dtype != v.dtype:
                    return False
            elif self.meta.get(k, None) is not v and self.meta.get(k, None) != v:
                return False
        return True

    def capture(self, mb: Dict[str, object]) -> None:
        self.static = {k: v.detach().clone() for k, v in mb.items() if torch.is_tensor(v)}
        self.meta = {k: v for k, v in mb.items() if not torch.is_tensor(v)}
        ftps = _fused_tp_objects()
        if ftps and os.environ.get("NXDT_CUDA_GRAPH_WITH_FUSED_TP", "0") != "1":
            # Measured on 2 B200s: replaying the fused compute+collective kernels from a graph desynchronises the
            # ranks' protocol epochs (deadlock caught by the in-kernel watchdog), while the NCCL-only graph replays
            # correctly.  Until that is understood the combination is refused (the caller falls back to eager).
            raise RuntimeError("CUDA-graph replay is not yet supported together with the fused TP kernels "
                               "(set NXDT_CUDA_GRAPH_WITH_FUSED_TP=1 to try anyway)")
        torch.cuda.synchronize()
        from .. import ops
        launches0 = ops.native_launch_count()
        g = torch.cuda.CUDAGraph()
        # thread_local: the NCCL watchdog thread keeps polling its own events while this thread captures
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            for f in ftps:
                f.graph_boundary()
            loss = self.fn({**self.static, **self.meta})
            self.loss = loss.detach()
            for f in ftps:
                f.graph_boundary()
        self.graph = g
        # kernels of this package recorded in the graph: capture does not execute them, every replay does
        self.native_launches = ops.native_launch_count() - launches0
        ops._count(-self.native_launches)
        torch.cuda.synchronize()

    def run(self, mb: Dict[str, object]) -> torch.Tensor:
        for k, s in self.static.items():
            s.copy_(mb[k], non_blocking=True)
        self.graph.replay()
        from .. import ops
        ops._count(self.native_launches)
        return self.loss
