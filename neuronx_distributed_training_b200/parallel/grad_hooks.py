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


# ---- read-modify-write updates of main_grad when micro-batches are issued on several CUDA streams --------------------
# (element-wise atomic writers -- the TMA reduce-add epilogues, the embedding scatter-add -- need no ordering)
_STATE = {"multi_stream": False, "last_event": None}


def set_multi_stream(flag: bool) -> None:
    _STATE["multi_stream"] = bool(flag)
    if not flag:
        _STATE["last_event"] = None


def multi_stream() -> bool:
    return _STATE["multi_stream"]


def ordered_main_grad_update(device, fn) -> None:
    """Run ``fn()`` (an in-place update of some ``main_grad``) ordered after the previous such update, whichever stream
    that ran on.  No-op wrapper unless the multi-stream mode is on."""
    import torch
    if not _STATE["multi_stream"] or device.type != "cuda":
        fn()
        return
    cur = torch.cuda.current_stream(device)
    prev = _STATE["last_event"]
    if prev is not None:
        cur.wait_event(prev)
    fn()
    ev = torch.cuda.Event()
    ev.record(cur)
    _STATE["last_event"] = ev
