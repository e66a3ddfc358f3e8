"""Training loop runtime.

Capability parity with the reference's Lightning trainer stack (SURVEY §2.1 #5-#7, §3.3, §3.5;
reference lightning_modules/nlp_overrides.py): ``max_steps`` / ``max_epochs`` / ``max_time``,
``val_check_interval``, ``limit_val_batches``, ``log_every_n_steps``, checkpoint callback, full state
restore (model, optimizer shards, LR schedule, loop progress, consumed samples), signal handling.

What is deliberately different (B200-first): there is no lazy-graph cut / step closure machinery; a step
is eager CUDA work on streams, metrics are read back only on logging steps, the step time reported is
measured with CUDA events on the device, and a SIGTERM/SIGUSR1 triggers an emergency checkpoint at the
next step boundary instead of relying on relaunch-only recovery.
"""
from __future__ import annotations

import logging
import os
import signal
import time
from typing import Any, Dict, Optional

import torch

from ..checkpoint import io as ckpt_io
from ..parallel import state as ps
from ..utils.exp_manager import ExpManager, consumed_samples_from_path, step_from_path
from ..utils.throughput import DeviceStepTimer
from ..utils.profiling import StepProfiler, nvtx_range

log = logging.getLogger("nxdt")


def _parse_max_time(v) -> Optional[float]:
    """'DD:HH:MM:SS' (PTL style) or seconds."""
    if v is None:
        return None
    if isinstance(v, (int, float)):
        return float(v)
    parts = [int(p) for p in str(v).split(":")]
    while len(parts) < 4:
        parts.insert(0, 0)
    d, h, m, s = parts
    return float(((d * 24 + h) * 60 + m) * 60 + s)


class Trainer:
    def __init__(self, devices: int = 1, num_nodes: int = 1, max_epochs: int = -1, max_steps: int = -1,
                 log_every_n_steps: int = 10, val_check_interval: Any = None, check_val_every_n_epoch: Any = None,
                 num_sanity_val_steps: int = 0, limit_val_batches: Any = 1.0, limit_test_batches: Any = 1.0,
                 gradient_clip_val: Optional[float] = None, max_time: Any = None, accumulate_grad_batches: int = 1,
                 enable_progress_bar: bool = False, profiler: Any = None, callbacks: Optional[list] = None, **ignored):
        if accumulate_grad_batches not in (None, 1):
            raise ValueError("accumulate_grad_batches>1 is not supported; gradient accumulation is expressed "
                             "through global_batch_size / micro_batch_size")
        if ignored.get("gradient_clip_algorithm") not in (None, "norm"):
            raise ValueError("only norm-based gradient clipping is supported")
        self.devices, self.num_nodes = devices, num_nodes
        self.max_epochs, self.max_steps = max_epochs, int(max_steps if max_steps is not None else -1)
        self.log_every_n_steps = max(1, int(log_every_n_steps or 1))
        self.val_check_interval = val_check_interval
        self.limit_val_batches = limit_val_batches
        self.limit_test_batches = limit_test_batches
        self.num_sanity_val_steps = int(num_sanity_val_steps or 0)
        self.gradient_clip_val = gradient_clip_val
        self.max_time = _parse_max_time(max_time)
        # user callbacks: objects with any of the PTL-style hooks on_fit_start / on_train_start / on_train_batch_end /
        # on_validation_end / on_save_checkpoint / on_train_end (each receives the trainer first)
        self.callbacks = list(callbacks or [])
        self.enable_progress_bar = bool(enable_progress_bar)
        self.ignored_args = ignored           # Neuron-only knobs (lnc, sequential_move_factor, ...)
        self.profiler_cfg = profiler          # {start_step, end_step, dir} -> torch.profiler trace of those steps
        self.profiler: Optional[StepProfiler] = None
        self.ema = None
        self.global_step = 0
        self.current_epoch = 0
        self.steps_since_resume = 0
        self.exp_manager: Optional[ExpManager] = None
        self.log_dir: Optional[str] = None
        self.model = None
        self.datamodule = None
        self.should_stop = False
        self._emergency = False
        self._sig_emergency = False
        self.logged_metrics: Dict[str, float] = {}
        self.callback_metrics: Dict[str, float] = {}
        self.step_timer = DeviceStepTimer()
        self.last_step_ms = []

    def _call(self, hook: str, *args) -> None:
        for cb in self.callbacks:
            fn = getattr(cb, hook, None)
            if fn is not None:
                fn(self, *args)

    # ---- stop decisions are collective -------------------------------------------------------------------
    def _agree(self, *flags: bool):
        """MAX-reduce local stop flags over all ranks: a signal lands on ONE rank and every rank reads its own
        clock, but leaving the step loop (and entering the checkpoint barrier) must happen on the same step
        everywhere, otherwise the remaining ranks hang in the next step's collectives."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return tuple(bool(f) for f in flags)
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([int(bool(f)) for f in flags], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return tuple(bool(v) for v in t.tolist())

    # ---- signals ---------------------------------------------------------------------------------------
    def _install_signal_handlers(self):
        def handler(signum, frame):
            log.warning("signal %s received: emergency checkpoint at next step boundary, then exit", signum)
            self._sig_emergency = True      # acted on at the next step boundary, after all ranks have agreed
        for sig in (signal.SIGTERM, signal.SIGUSR1):
            try:
                signal.signal(sig, handler)
            except (ValueError, OSError):
                pass

    # ---- checkpoint ------------------------------------------------------------------------------------
    def dump_checkpoint(self) -> Dict[str, Any]:
        return {
            "epoch": self.current_epoch, "global_step": self.global_step,
            "pytorch-lightning_version": "2.5.0",
            "loops": {"fit_loop": {"global_step": self.global_step, "epoch": self.current_epoch,
                                   "consumed_samples": self.datamodule.consumed_samples if self.datamodule else 0}},
            "lr_schedulers": [self.model.lr_scheduler.state_dict()] if self.model.lr_scheduler is not None else [],
            "callbacks": {}, "hyper_parameters": {"cfg": self.model.cfg.to_dict()},
            "datamodule": self.datamodule.state_dict() if self.datamodule is not None else {},
        }

    def save_checkpoint(self, filepath: str):
        em = self.exp_manager.cfg if self.exp_manager else None
        directory, tag = os.path.dirname(filepath), os.path.basename(filepath)
        # strip the "-vN" version suffix PTL appends (reference nlp_overrides.py:572-578)
        user_content = self.dump_checkpoint()
        model_sd = self.model.state_dict()
        if em is not None and em.save_bf16:
            # exp_manager.save_bf16: store the model weights in bf16 regardless of the training dtype
            model_sd = {k: (v.to(torch.bfloat16) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in model_sd.items()}
        extra = self.model.extra_checkpoint_files() if hasattr(self.model, "extra_checkpoint_files") else None
        if extra:
            user_content[ckpt_io.EXTRA_FILES_KEY] = extra
        opt = self.model.optimizer
        if em is not None and em.checkpoint_callback_params.use_master_weights_in_ckpt and opt.master is not None:
            user_content["master_weights_in_ckpt"] = True
        opt_sd = opt.state_dict() if opt is not None else None
        ckpt_io.save_checkpoint(directory, tag, model=model_sd, optimizer=opt_sd, user_content=user_content,
                                use_xser=em.save_xser if em else True,
                                async_save=em.async_checkpointing if em else False,
                                zero1_optimizer=bool(opt is not None and opt.zero1))
        if ps.get_rank() == 0:
            log.info("saved checkpoint %s", filepath)

    def restore(self, ckpt_path: str, weight_init_only: bool = False):
        directory, tag = os.path.dirname(ckpt_path.rstrip("/")), os.path.basename(ckpt_path.rstrip("/"))
        m, o, uc = ckpt_io.load_checkpoint(directory, tag, model=True, optimizer=not weight_init_only)
        self.model.load_state_dict(m)
        opt = self.model.optimizer
        if weight_init_only or o is None:
            opt.sync_master_from_params()
            log.info("loaded weights only from %s", ckpt_path)
            return
        if hasattr(opt, "shard_lo") and not ckpt_io._is_remote(directory):
            # ZeRO-1 shards written with another data-parallel size are re-assembled for this rank's slice
            from ..checkpoint.reshard import load_optimizer_state_any_dp
            if load_optimizer_state_any_dp(opt, o, directory, tag, ps.get_tensor_model_parallel_rank(),
                                           ps.get_pipeline_model_parallel_rank()):
                log.info("optimizer state re-sharded for data-parallel size %d", opt.dp_size)
        else:
            opt.load_state_dict(o)
        if uc is not None:
            if "pytorch-lightning_version" not in uc:
                uc["pytorch-lightning_version"] = (self.exp_manager.cfg.ckpt_ptl_version if self.exp_manager else None) or "2.5.0"
            self.global_step = int(uc.get("global_step", step_from_path(ckpt_path)))
            self.current_epoch = int(uc.get("epoch", 0))
            scheds = uc.get("lr_schedulers") or []
            if scheds and self.model.lr_scheduler is not None:
                self.model.lr_scheduler.load_state_dict(scheds[0])
            if self.datamodule is not None and uc.get("datamodule"):
                self.datamodule.load_state_dict(uc["datamodule"])
        else:
            self.global_step = step_from_path(ckpt_path)
        if self.datamodule is not None and consumed_samples_from_path(ckpt_path) > 0:
            self.datamodule.load_state_dict({"consumed_samples": consumed_samples_from_path(ckpt_path)})
        log.info("resumed from %s at global_step=%d consumed_samples=%d", ckpt_path, self.global_step,
                 self.datamodule.consumed_samples if self.datamodule else -1)

    # ---- validation --------------------------------------------------------------------------------------
    def _limit(self, limit, n_total: Optional[int]) -> int:
        if limit is None:
            return n_total if n_total is not None else 1 << 30
        if isinstance(limit, float) and limit <= 1.0:
            return int((n_total or 0) * limit)
        return int(limit)

    def validate(self) -> Optional[float]:
        loader = self.datamodule.val_dataloader() if self.datamodule is not None else None
        if loader is None:
            return None
        try:
            n_total = len(loader)
        except TypeError:
            n_total = None
        n = self._limit(self.limit_val_batches, n_total)
        if n <= 0:
            return None
        losses = []
        if self.ema is not None and (self.exp_manager.cfg.ema or {}).get("validate_original_weights", False) is False:
            self.ema.swap_in()
        try:
            for i, batch in enumerate(loader):
                if i >= n:
                    break
                with nvtx_range("validation_step"):
                    losses.append(self.model.validation_step(batch, i))
        finally:
            if self.ema is not None:
                self.ema.swap_out()
        if not losses:
            return None
        val = float(torch.stack([l.detach().float() for l in losses]).mean())
        self.callback_metrics["val_loss"] = val
        if self.exp_manager:
            self.exp_manager.log_metrics({"val_loss": val}, self.global_step)
        log.info("validation: val_loss=%.5f (%d batches)", val, len(losses))
        return val

    # ---- test / predict (the evaluation and prediction loops of the reference trainer, nlp_overrides.py:264-533) -----
    def _attach(self, model, datamodule, ckpt_path: Optional[str]):
        """Set a model / data module up outside ``fit`` (weights from ``ckpt_path`` when given)."""
        self.model, self.datamodule = model, datamodule
        model.trainer = self
        if datamodule is not None:
            datamodule.setup(ckpt_path)
        model.setup("test")
        if getattr(model, "optimizer", None) is None:
            model.configure_optimizers(self.max_steps if self.max_steps > 0 else None)
        if ckpt_path:
            self.restore(ckpt_path, weight_init_only=True)

    def test(self, model=None, datamodule=None, ckpt_path: Optional[str] = None) -> Optional[float]:
        """Mean loss over ``test_dataloader`` (``limit_test_batches`` applies); logged as ``test_loss``."""
        if model is not None:
            self._attach(model, datamodule, ckpt_path)
        loader = self.datamodule.test_dataloader() if self.datamodule is not None and hasattr(self.datamodule, "test_dataloader") else None
        if loader is None:
            return None
        try:
            n_total = len(loader)
        except TypeError:
            n_total = None
        n = self._limit(self.limit_test_batches, n_total)
        losses = []
        for i, batch in enumerate(loader):
            if i >= n:
                break
            with nvtx_range("test_step"):
                losses.append(self.model.validation_step(batch, i))
        if not losses:
            return None
        val = float(torch.stack([l.detach().float() for l in losses]).mean())
        self.callback_metrics["test_loss"] = val
        if self.exp_manager:
            self.exp_manager.log_metrics({"test_loss": val}, self.global_step)
        log.info("test: test_loss=%.5f (%d batches)", val, len(losses))
        return val

    def predict(self, model=None, datamodule=None, ckpt_path: Optional[str] = None, dataloader=None) -> list:
        """Runs ``model.predict_step(batch, idx)`` over a dataloader (``predict_dataloader`` of the data module, else its
        test / validation loader) and returns the list of per-batch outputs."""
        if model is not None:
            self._attach(model, datamodule, ckpt_path)
        loader = dataloader
        if loader is None and self.datamodule is not None:
            for name in ("predict_dataloader", "test_dataloader", "val_dataloader"):
                fn = getattr(self.datamodule, name, None)
                loader = fn() if fn is not None else None
                if loader is not None:
                    break
        if loader is None:
            return []
        outs = []
        for i, batch in enumerate(loader):
            with nvtx_range("predict_step"):
                outs.append(self.model.predict_step(batch, i))
        return outs

    # ---- fit -----------------------------------------------------------------------------------------------
    def fit(self, model, datamodule, ckpt_path: Optional[str] = None):
        self.model, self.datamodule = model, datamodule
        model.trainer = self
        self._install_signal_handlers()
        if ckpt_path is None and self.exp_manager is not None:
            ckpt_path = self.exp_manager.resume_from_checkpoint
        datamodule.setup(ckpt_path)
        model.setup("fit")
        model.configure_optimizers(self.max_steps if self.max_steps > 0 else None)
        weight_init_only = bool(model.cfg.model.get("weight_init_only", False))
        if ckpt_path:
            self.restore(ckpt_path, weight_init_only=weight_init_only and
                         not os.path.dirname(ckpt_path.rstrip("/")).startswith(str(self.log_dir or "\0")))
        self._call("on_fit_start", model)
        model.on_train_start(self)
        self._call("on_train_start", model)
        if self.profiler_cfg:
            pc = dict(self.profiler_cfg) if not hasattr(self.profiler_cfg, "to_dict") else self.profiler_cfg.to_dict()
            self.profiler = StepProfiler(pc.get("start_step", 3), pc.get("end_step", 5),
                                         pc.get("dir", os.path.join(self.log_dir or ".", "profiler")), ps.get_rank())
        ema_cfg = self.exp_manager.cfg.ema if self.exp_manager is not None else None
        if ema_cfg and ema_cfg.get("enable", False):
            from ..utils.ema import EMA
            self.ema = EMA(model.optimizer, float(ema_cfg.get("decay", 0.999)), int(ema_cfg.get("every_n_steps", 1)))
        for i in range(self.num_sanity_val_steps and 1):
            self.validate()
        t_start = time.time()
        cb = self.exp_manager.checkpoint_callback if self.exp_manager else None
        emc = self.exp_manager.cfg if self.exp_manager else None
        if self.max_time is None and emc is not None and emc.max_time_per_run:
            self.max_time = _parse_max_time(emc.max_time_per_run)       # NeMo's per-run wall-clock budget
        first_step_pending = bool(emc is not None and emc.enable_recovery_time_instrumentation)
        # hang detection: top-level `neuron_rt_exec_timeout` (the reference's runtime timeout knob) or
        # NXDT_STEP_TIMEOUT, in seconds; 0 / unset = off
        timeout = float(os.environ.get("NXDT_STEP_TIMEOUT", model.cfg.get("neuron_rt_exec_timeout", 0) or 0) or 0)
        self.watchdog = None
        if timeout > 0:
            from ..utils.watchdog import StepWatchdog
            self.watchdog = StepWatchdog(timeout, os.path.join(self.log_dir or ".", f"nemo_error_log_rank{ps.get_rank()}.txt")).start()
        try:
            while not self.should_stop:
                loader = datamodule.train_dataloader()
                n_batches = 0
                for batch_idx, batch in enumerate(loader):
                    if self.max_steps > 0 and self.global_step >= self.max_steps:
                        self.should_stop = True
                        break
                    if self.profiler is not None:
                        self.profiler.step(self.global_step)
                    self.step_timer.start()
                    with nvtx_range(f"train_step_{self.global_step}"):
                        loss = model.training_step(batch, batch_idx)
                    self.step_timer.stop()
                    if self.ema is not None:
                        self.ema.update(self.global_step + 1)
                    n_batches += 1
                    if first_step_pending:
                        # time from process start-up (exp manager creation) to the end of the first step after a (re)start
                        if torch.cuda.is_available():
                            torch.cuda.synchronize()
                        rec = time.time() - self.exp_manager.t_created
                        log.info("recovery time: first training step finished %.2f s after start-up (resumed=%s)", rec, bool(ckpt_path))
                        self.callback_metrics["recovery_time_s"] = rec
                        first_step_pending = False
                    if self.watchdog is not None:
                        self.watchdog.pet()
                    self.global_step += 1
                    self.steps_since_resume += 1
                    datamodule.consumed_samples = datamodule.compute_consumed_samples(self.steps_since_resume)
                    if self.global_step % self.log_every_n_steps == 0:
                        self._log(loss)
                    self._call("on_train_batch_end", model, loss, batch, batch_idx)
                    vci = self.val_check_interval
                    if vci and isinstance(vci, int) and self.global_step % vci == 0:
                        self.validate()
                        self._call("on_validation_end", model)
                    if cb is not None and cb.should_save(self.global_step):
                        cb.save(self, self._ckpt_metrics())
                        self._call("on_save_checkpoint", model)
                    if self.watchdog is not None:
                        self.watchdog.pet()          # validation / checkpointing are not part of the step budget
                    timed_out = self.max_time is not None and time.time() - t_start > self.max_time
                    emergency, timed_out, stop = self._agree(self._sig_emergency, timed_out, self.should_stop)
                    if timed_out:
                        log.info("max_time reached")
                    if emergency:
                        self._emergency = True
                    self.should_stop = stop or emergency or timed_out
                    if self.should_stop:
                        break
                self.current_epoch += 1
                if n_batches == 0:
                    if not self.should_stop and (self.max_steps <= 0 or self.global_step < self.max_steps):
                        log.warning("the training dataloader yielded no batch in epoch %d (dataset of %s samples, global "
                                    "batch %s): stopping at step %d before max_steps", self.current_epoch - 1,
                                    getattr(getattr(loader, "dataset", None), "__len__", lambda: "?")(),
                                    getattr(datamodule, "global_batch_size", "?"), self.global_step)
                    break
                if self.max_epochs and self.max_epochs > 0 and self.current_epoch >= self.max_epochs:
                    break
                if self.max_steps > 0 and self.global_step >= self.max_steps:
                    break
            if cb is not None and (self._emergency or cb.p.save_last) and self.global_step != cb.last_saved_step \
                    and self.global_step > 0:
                cb.save(self, self._ckpt_metrics(), last=not self._emergency and cb.p.save_last)
        finally:
            if self.watchdog is not None:
                self.watchdog.stop()
            model.on_train_end(self)
            self._call("on_train_end", model)
            ckpt_io.finalize_checkpoint()
            if self.exp_manager:
                self.exp_manager.close()
        return self

    def _ckpt_metrics(self) -> Dict[str, Any]:
        return {"step": self.global_step, "consumed_samples": float(self.datamodule.consumed_samples),
                "epoch": self.current_epoch, **self.callback_metrics}

    def _log(self, loss: torch.Tensor):
        m = self.model.collect_metrics(loss, self.global_step, self.datamodule.consumed_samples)
        ms = self.step_timer.elapsed_ms()
        if ms:
            m["step_time_ms_device"] = sum(ms) / len(ms)
            self.last_step_ms = ms
        self.step_timer.reset()
        self.logged_metrics = m
        self.callback_metrics.update({"reduced_train_loss": m["reduced_train_loss"]})
        if self.exp_manager:
            self.exp_manager.log_metrics(m, self.global_step)
        if self.enable_progress_bar and ps.get_rank() == 0 and self.max_steps > 0:
            done = int(30 * self.global_step / self.max_steps)
            print(f"\r[{'#' * done}{'.' * (30 - done)}] {self.global_step}/{self.max_steps} loss {m['reduced_train_loss']:.4f}",
                  end="" if self.global_step < self.max_steps else "\n", flush=True)
        if ps.get_rank() == 0:
            log.info("step %d | %s", self.global_step,
                     " | ".join(f"{k}={v:.6g}" for k, v in m.items() if k != "global_step"))


NLPTrainer = Trainer
