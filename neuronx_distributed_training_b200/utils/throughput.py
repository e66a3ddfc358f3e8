"""Throughput meters.

The reference logs a 10-step moving average of sequences/s on the host clock (reference
utils/utils.py:52-77).  We keep that metric name/semantics for the logs and add a device-timed meter
(CUDA events on the compute stream) which is the number the benchmark reports.
"""
from __future__ import annotations

import time
from collections import deque
from typing import Deque, Optional

import torch


class Throughput:
    def __init__(self, batch_size: int, world_size_dp: int, grad_accum: int, moving_avg_window_size: int = 10):
        self.seqs_per_iteration = batch_size * world_size_dp * grad_accum
        self.window: Deque[float] = deque(maxlen=moving_avg_window_size)
        self.start = time.time()
        self.throughput_peak = 0.0
        self.throughput_sum = 0.0
        self.n = 0

    def get_throughput(self) -> float:
        now = time.time()
        self.window.append(now - self.start)
        self.start = now
        tput = len(self.window) * self.seqs_per_iteration / max(sum(self.window), 1e-9)
        self.n += 1
        self.throughput_sum += tput
        self.throughput_peak = max(self.throughput_peak, tput)
        return tput


class DeviceStepTimer:
    """CUDA-event step timer (falls back to the host clock on CPU)."""

    def __init__(self, device: Optional[torch.device] = None):
        self.cuda = torch.cuda.is_available() and (device is None or torch.device(device).type == "cuda")
        self._start = None
        self._pairs = []
        self._t0 = 0.0
        self.host_ms = []

    def start(self):
        if self.cuda:
            self._start = torch.cuda.Event(enable_timing=True)
            self._start.record()
        self._t0 = time.perf_counter()

    def stop(self):
        if self.cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._pairs.append((self._start, e))
        self.host_ms.append((time.perf_counter() - self._t0) * 1e3)

    def elapsed_ms(self):
        """list of per-step device times (synchronises)."""
        if not self.cuda:
            return list(self.host_ms)
        torch.cuda.synchronize()
        return [s.elapsed_time(e) for s, e in self._pairs]

    def reset(self):
        self._pairs.clear()
        self.host_ms.clear()
