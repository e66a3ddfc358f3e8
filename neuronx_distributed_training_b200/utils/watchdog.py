"""Hang detection for a training rank.

The reference bounds a device execution with ``neuron_rt_exec_timeout`` (examples/training_orchestrator.py:45) and
relies on relaunch + auto-resume for recovery.  Here a daemon thread watches the step counter: if no step completes
within ``timeout_s`` it dumps every Python thread's stack (``faulthandler``) to the rank's error log, so a stuck
collective or a deadlocked peer handshake is visible in the logs, and -- unless ``abort=False`` -- exits the process
with code 124.  Under ``torchrun --max-restarts`` that tears the job down and relaunches it, and the experiment
manager resumes from the newest complete checkpoint.  (CUDA kernels that spin on a peer additionally trap after ~20 s
on the device, see ``ops/csrc/sm100_ptx.cuh``; NCCL has its own watchdog.)
"""
from __future__ import annotations

import faulthandler
import logging
import os
import sys
import threading
import time
from typing import Callable, Optional

log = logging.getLogger("nxdt")


class StepWatchdog:
    def __init__(self, timeout_s: float, dump_path: Optional[str] = None, abort: bool = True,
                 on_timeout: Optional[Callable[[float], None]] = None, poll_s: Optional[float] = None):
        self.timeout_s = float(timeout_s)
        self.dump_path = dump_path
        self.abort = abort
        self.on_timeout = on_timeout
        self.poll_s = poll_s if poll_s is not None else max(0.05, min(5.0, self.timeout_s / 10.0))
        self._last = time.monotonic()
        self._paused = False
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self.fired = False

    def start(self) -> "StepWatchdog":
        if self.timeout_s <= 0 or self._thread is not None:
            return self
        self._last = time.monotonic()
        self._thread = threading.Thread(target=self._run, name="nxdt-step-watchdog", daemon=True)
        self._thread.start()
        return self

    def pet(self) -> None:
        """Call once per completed step (and around long legitimate pauses such as checkpointing)."""
        self._last = time.monotonic()

    def pause(self) -> None:
        self._paused = True

    def resume(self) -> None:
        self._last = time.monotonic()
        self._paused = False

    def stop(self) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2.0)
            self._thread = None

    def _run(self) -> None:
        while not self._stop.wait(self.poll_s):
            if self._paused:
                continue
            idle = time.monotonic() - self._last
            if idle > self.timeout_s:
                self.fired = True
                self._report(idle)
                if self.on_timeout is not None:
                    self.on_timeout(idle)
                if self.abort:
                    os._exit(124)
                self._last = time.monotonic()

    def _report(self, idle: float) -> None:
        msg = f"step watchdog: no training step completed for {idle:.1f} s (limit {self.timeout_s:.1f} s); thread stacks follow"
        try:
            log.error(msg)
        except Exception:  # noqa
            pass
        try:
            if self.dump_path:
                os.makedirs(os.path.dirname(self.dump_path) or ".", exist_ok=True)
                with open(self.dump_path, "a") as f:
                    f.write(msg + "\n")
                    f.flush()
                    faulthandler.dump_traceback(file=f, all_threads=True)
            else:
                faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        except Exception:  # noqa
            pass
