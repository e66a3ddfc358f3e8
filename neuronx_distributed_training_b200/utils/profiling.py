"""Tracing / profiling helpers: NVTX ranges per phase, an optional ``torch.profiler`` window, and the
device-timed step meter.  The reference only forwards PTL's ``profiler`` argument and a host-side TimingCallback; on B200
the useful artefacts are NVTX ranges (visible in nsys / ncu) and CUDA-event timings."""
from __future__ import annotations

import contextlib
import os

import torch

_NVTX = os.environ.get("NXDT_NVTX", "0") == "1"


def enable_nvtx(flag: bool = True) -> None:
    global _NVTX
    _NVTX = flag


@contextlib.contextmanager
def nvtx_range(name: str):
    """No-op unless NXDT_NVTX=1 (or enable_nvtx()) and CUDA is present."""
    if _NVTX and torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield


class StepProfiler:
    """``trainer.profiler: {start_step, end_step, dir}`` -> a torch.profiler trace (Chrome / TensorBoard format) of the
    chosen steps on rank 0."""

    def __init__(self, start_step: int, end_step: int, out_dir: str, rank: int = 0):
        self.start, self.end, self.out_dir, self.rank = int(start_step), int(end_step), out_dir, rank
        self.prof = None

    def step(self, global_step: int) -> None:
        if self.rank != 0:
            return
        if global_step == self.start and self.prof is None:
            acts = [torch.profiler.ProfilerActivity.CPU]
            if torch.cuda.is_available():
                acts.append(torch.profiler.ProfilerActivity.CUDA)
            os.makedirs(self.out_dir, exist_ok=True)
            self.prof = torch.profiler.profile(activities=acts, record_shapes=False)
            self.prof.__enter__()
        elif global_step == self.end and self.prof is not None:
            self.prof.__exit__(None, None, None)
            self.prof.export_chrome_trace(os.path.join(self.out_dir, f"trace_steps_{self.start}_{self.end}.json"))
            self.prof = None
