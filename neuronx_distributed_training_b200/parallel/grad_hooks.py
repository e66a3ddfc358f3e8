"""Notification point for code that writes a weight gradient straight into ``param.main_grad`` (wgrad GEMM epilogues,
the fused LM-head / cross-entropy, the embedding scatter-add, grouped expert GEMMs).  The optimizer may hang a callback on
the parameter to learn that its gradient for the current micro-batch is complete (used to start the data-parallel
reduction of a gradient bucket while backward is still running)."""
from __future__ import annotations


def mark_main_grad_written(param) -> None:
    param.grad_added_to_main_grad = True
    cb = getattr(param, "_grad_ready_cb", None)
    if cb is not None:
        cb(param)
