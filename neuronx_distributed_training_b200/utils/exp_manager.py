"""Experiment manager: log directories, loggers, auto-resume, checkpoint callback, step timing.

Capability parity with the reference's ``exp_manager`` (reference
utils/exp_manager.py:81-579): ``<exp_dir or ./nemo_experiments>/<name>/<version>/{checkpoints,...}``,
per-rank ``nemo_log_globalrank-*_localrank-*.txt``, ``cmd-args.log``, ``nemo_error_log.txt``,
TensorBoard event files, newest-``*.ckpt`` auto-resume with old logs moved to ``run_N/``, top-k /
every-n-steps / time-interval checkpointing with the ``{step}-{consumed_samples}`` filename template.
"""
from __future__ import annotations

import glob
import logging
import os
import re
import shutil
import sys
import time
from dataclasses import dataclass, field, fields
from typing import Any, Dict, List, Optional

import torch
import torch.distributed as dist

from ..config import Config
from ..parallel import state as ps
from ..checkpoint import io as ckpt_io

log = logging.getLogger("nxdt")


@dataclass
class CallbackParams:
    filepath: Optional[str] = None
    dirpath: Optional[str] = None
    filename: Optional[str] = None
    monitor: str = "step"
    verbose: bool = True
    save_last: bool = True
    save_top_k: int = 3
    save_weights_only: bool = False
    mode: str = "max"
    every_n_epochs: Optional[int] = None
    every_n_train_steps: Optional[int] = None
    train_time_interval: Optional[Any] = None
    prefix: Optional[str] = None
    postfix: str = ".nemo"
    save_best_model: bool = False
    always_save_nemo: bool = False
    save_nemo_on_train_end: bool = False
    model_parallel_size: Optional[int] = None
    save_on_train_epoch_end: bool = False
    use_master_weights_in_ckpt: bool = False
    enable_version_counter: bool = True
    save_lora_base: bool = False
    save_lora_config_adapter: bool = True
    merge_lora: bool = False


@dataclass
class ExpManagerConfig:
    explicit_log_dir: Optional[str] = None
    exp_dir: Optional[str] = None
    name: Optional[str] = None
    version: Optional[str] = None
    use_datetime_version: bool = True
    resume_if_exists: bool = False
    resume_past_end: bool = False
    resume_ignore_no_checkpoint: bool = False
    resume_from_checkpoint: Optional[str] = None
    create_tensorboard_logger: bool = True
    summary_writer_kwargs: Optional[dict] = None
    create_wandb_logger: bool = False
    wandb_logger_kwargs: Optional[dict] = None
    create_mlflow_logger: bool = False
    mlflow_logger_kwargs: Optional[dict] = None
    create_checkpoint_callback: bool = True
    checkpoint_callback_params: CallbackParams = field(default_factory=CallbackParams)
    files_to_copy: Optional[list] = None
    log_step_timing: bool = True
    step_timing_kwargs: Optional[dict] = None
    log_local_rank_0_only: bool = False
    log_global_rank_0_only: bool = False
    log_parameter_norm: bool = True       # reference defaults (utils/exp_manager.py:52-53)
    log_gradient_norm: bool = True
    enable_recovery_time_instrumentation: bool = False
    save_xser: bool = True
    load_xser: bool = True
    save_bf16: bool = False
    async_checkpointing: bool = False
    ckpt_ptl_version: Optional[str] = None
    ema: Optional[dict] = None
    max_time_per_run: Optional[str] = None
    seconds_to_sleep: float = 0.0

    @staticmethod
    def from_config(node: Optional[Config]) -> "ExpManagerConfig":
        d = node.to_dict() if node is not None else {}
        known = {f.name for f in fields(ExpManagerConfig)}
        unknown = [k for k in d if k not in known]
        if unknown:
            raise ValueError(f"unknown exp_manager keys: {unknown}")
        cb = d.pop("checkpoint_callback_params", None) or {}
        cb_known = {f.name for f in fields(CallbackParams)}
        bad = [k for k in cb if k not in cb_known]
        if bad:
            raise ValueError(f"unknown checkpoint_callback_params keys: {bad}")
        return ExpManagerConfig(checkpoint_callback_params=CallbackParams(**cb), **d)


# ---- loggers -------------------------------------------------------------------------------------------
class TensorBoardLogger:
    def __init__(self, log_dir: str, **kw):
        self.writer = None
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.writer = SummaryWriter(log_dir=log_dir, **kw)
        except Exception as e:  # tensorboard missing: degrade to a jsonl scalar log
            log.warning("TensorBoard unavailable (%s); writing scalars to metrics.jsonl only", e)

    def log_metrics(self, metrics: Dict[str, float], step: int):
        if self.writer is not None:
            for k, v in metrics.items():
                self.writer.add_scalar(k, v, step)

    def close(self):
        if self.writer is not None:
            self.writer.flush()
            self.writer.close()


class WandbLogger:
    """``exp_manager.create_wandb_logger`` + ``wandb_logger_kwargs`` (name / project / ...); needs the wandb package."""

    def __init__(self, log_dir: str, **kw):
        self.run = None
        try:
            import wandb
            self.run = wandb.init(dir=log_dir, **kw)
        except Exception as e:  # package missing or no network: keep training, scalars still go to metrics.jsonl
            log.warning("wandb logger unavailable (%s); skipped", e)

    def log_metrics(self, metrics: Dict[str, float], step: int):
        if self.run is not None:
            self.run.log(dict(metrics), step=step)

    def close(self):
        if self.run is not None:
            self.run.finish()


class MLflowLogger:
    """``exp_manager.create_mlflow_logger`` + ``mlflow_logger_kwargs`` (experiment_name / tracking_uri / run_name ...)."""

    def __init__(self, log_dir: str, experiment_name: Optional[str] = None, tracking_uri: Optional[str] = None,
                 run_name: Optional[str] = None, tags: Optional[dict] = None, **unused):
        self.mlflow = None
        try:
            import mlflow
            mlflow.set_tracking_uri(tracking_uri or "file:" + os.path.join(log_dir, "mlruns"))
            if experiment_name:
                mlflow.set_experiment(experiment_name)
            mlflow.start_run(run_name=run_name, tags=tags)
            self.mlflow = mlflow
        except Exception as e:
            log.warning("mlflow logger unavailable (%s); skipped", e)

    def log_metrics(self, metrics: Dict[str, float], step: int):
        if self.mlflow is not None:
            self.mlflow.log_metrics({k.replace("/", "_"): float(v) for k, v in metrics.items()}, step=step)

    def close(self):
        if self.mlflow is not None:
            self.mlflow.end_run()


class JsonlLogger:
    def __init__(self, path: str):
        self.f = open(path, "a")

    def log_metrics(self, metrics: Dict[str, float], step: int):
        import json
        self.f.write(json.dumps({"step": step, **metrics}) + "\n")
        self.f.flush()

    def close(self):
        self.f.close()


class TimingCallback:
    """Logs wall time of the named trainer hooks (reference exp_manager.py:64-78)."""

    def __init__(self):
        self.t: Dict[str, float] = {}
        self.last: Dict[str, float] = {}

    def start(self, name):
        self.t[name] = time.perf_counter()

    def stop(self, name):
        if name in self.t:
            self.last[name + "_timing"] = time.perf_counter() - self.t.pop(name)

    def metrics(self):
        return dict(self.last)


# ---- resume --------------------------------------------------------------------------------------------
_CKPT_RE = re.compile(r"consumed_samples=([0-9]+(?:\.[0-9]+)?)")
_STEP_RE = re.compile(r"step=([0-9]+)")


def consumed_samples_from_path(path: Optional[str]) -> int:
    """reference data/base.py:40-47: parsed from the checkpoint file name."""
    if not path:
        return 0
    m = _CKPT_RE.search(os.path.basename(path.rstrip("/")))
    return int(float(m.group(1))) if m else 0


def step_from_path(path: Optional[str]) -> int:
    if not path:
        return 0
    m = _STEP_RE.search(os.path.basename(path.rstrip("/")))
    return int(m.group(1)) if m else 0


def find_latest_checkpoint(ckpt_dir: str) -> Optional[str]:
    """Newest complete ``*.ckpt`` directory (by mtime), ignoring ones without a ``done`` marker."""
    cands = [p for p in glob.glob(os.path.join(ckpt_dir, "*.ckpt")) if os.path.exists(os.path.join(p, "done")) or os.path.isfile(p)]
    if not cands:
        return None
    return max(cands, key=os.path.getmtime)


def check_resume(cfg: ExpManagerConfig, log_dir: str) -> Optional[str]:
    if cfg.resume_from_checkpoint:
        return cfg.resume_from_checkpoint
    if not cfg.resume_if_exists:
        return None
    ckpt_dir = os.path.join(log_dir, "checkpoints")
    latest = find_latest_checkpoint(ckpt_dir) if os.path.isdir(ckpt_dir) else None
    if latest is None:
        if cfg.resume_ignore_no_checkpoint:
            log.info("no checkpoint under %s; starting from scratch", ckpt_dir)
            return None
        raise FileNotFoundError(f"resume_if_exists=True but no checkpoint under {ckpt_dir}")
    # move stale logs out of the way: run_0, run_1, ...
    if ps.get_rank() == 0:
        stale = [f for f in os.listdir(log_dir) if os.path.isfile(os.path.join(log_dir, f))]
        if stale:
            n = len(glob.glob(os.path.join(log_dir, "run_*")))
            dest = os.path.join(log_dir, f"run_{n}")
            os.makedirs(dest, exist_ok=True)
            for f in stale:
                shutil.move(os.path.join(log_dir, f), os.path.join(dest, f))
    return latest


# ---- checkpoint callback ------------------------------------------------------------------------------
class ModelCheckpoint:
    """top-k / every-n-steps / time-interval checkpointing (reference exp_manager.py:407-514)."""

    def __init__(self, dirpath: str, params: CallbackParams, save_xser: bool, async_save: bool):
        self.dirpath, self.p = dirpath, params
        self.save_xser, self.async_save = save_xser, async_save
        self.saved: List[tuple] = []   # (score, path)
        self.last_time = time.time()
        self.last_saved_step = -1
        self.best_model_path = ""
        os.makedirs(dirpath, exist_ok=True) if ps.get_rank() == 0 else None
        # adopt checkpoints that already exist (resume) so top-k pruning keeps counting them
        for pth in sorted(glob.glob(os.path.join(dirpath, "*.ckpt")), key=os.path.getmtime):
            self.saved.append((float(step_from_path(pth)), pth))

    def format_name(self, metrics: Dict[str, Any]) -> str:
        tmpl = self.p.filename or "{step}"
        def sub(m):
            k = m.group(1)
            v = metrics.get(k, 0)
            return f"{k}={v}"
        name = re.sub(r"\{(\w+)(?::[^}]*)?\}", sub, tmpl)
        return name + ".ckpt"

    def _interval_seconds(self) -> Optional[float]:
        t = self.p.train_time_interval
        if t is None:
            return None
        if isinstance(t, (int, float)):
            return float(t)
        if isinstance(t, dict):
            return float(t.get("seconds", 0)) + 60.0 * float(t.get("minutes", 0)) + 3600.0 * float(t.get("hours", 0))
        return None

    def should_save(self, global_step: int) -> bool:
        if global_step == self.last_saved_step:
            return False
        n = self.p.every_n_train_steps
        if n and n > 0 and global_step % n == 0:
            return True
        iv = self._interval_seconds()
        if iv is not None:
            decide = torch.tensor([1 if time.time() - self.last_time >= iv else 0])
            if dist.is_initialized() and dist.get_world_size() > 1:
                decide = decide.to("cuda" if dist.get_backend() == "nccl" else "cpu")
                dist.broadcast(decide, src=0)      # all ranks must agree (reference :476)
            return bool(decide.item())
        return False

    def save(self, trainer, metrics: Dict[str, Any], last: bool = False) -> str:
        name = self.format_name(metrics)
        if last:
            name = name[:-5] + "-last.ckpt"
        path = os.path.join(self.dirpath, name)
        trainer.save_checkpoint(path)
        self.last_time = time.time()
        self.last_saved_step = int(metrics.get("step", -1))
        score = float(metrics.get(self.p.monitor, metrics.get("step", 0)))
        if not last:
            self.saved.append((score, path))
            self.best_model_path = path
            self._prune()
        return path

    def _prune(self):
        k = self.p.save_top_k
        if k is None or k < 0:
            return
        rev = self.p.mode == "max"
        self.saved.sort(key=lambda t: t[0], reverse=rev)
        while len(self.saved) > max(k, 0):
            _, victim = self.saved.pop()
            ckpt_io.remove_checkpoint(victim, async_remove=self.async_save)


# ---- main entry ----------------------------------------------------------------------------------------
class ExpManager:
    def __init__(self, trainer, cfg_node: Optional[Config]):
        self.cfg = ExpManagerConfig.from_config(cfg_node)
        c = self.cfg
        rank = ps.get_rank() if ps.model_parallel_is_initialized() else int(os.environ.get("RANK", "0"))
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if c.explicit_log_dir:
            self.log_dir = str(c.explicit_log_dir)
        else:
            base = c.exp_dir or os.path.join(os.getcwd(), "nemo_experiments")
            name = c.name or "default"
            version = c.version or os.environ.get("NEMO_EXPM_VERSION")
            if version is None:
                version = "version_0" if (c.resume_if_exists or not c.use_datetime_version) else time.strftime("%Y-%m-%d_%H-%M-%S")
            self.log_dir = os.path.join(base, name, str(version))
        if rank == 0:
            os.makedirs(self.log_dir, exist_ok=True)
        if dist.is_initialized():
            dist.barrier()
        self.resume_from_checkpoint = check_resume(c, self.log_dir)
        self.loggers = []
        self._setup_text_logging(rank, local_rank)
        if rank == 0:
            with open(os.path.join(self.log_dir, "cmd-args.log"), "a") as f:
                f.write(" ".join(sys.argv) + "\n")
            self.loggers.append(JsonlLogger(os.path.join(self.log_dir, "metrics.jsonl")))
            for src in (c.files_to_copy or []):          # keep the launch inputs next to the logs (NeMo behaviour)
                try:
                    import shutil
                    shutil.copy(str(src), self.log_dir)
                except OSError as e:
                    log.warning("files_to_copy: could not copy %s (%s)", src, e)
            if c.create_tensorboard_logger:
                self.loggers.append(TensorBoardLogger(self.log_dir, **(c.summary_writer_kwargs or {})))
            if c.create_wandb_logger:
                self.loggers.append(WandbLogger(self.log_dir, **(c.wandb_logger_kwargs or {})))
            if c.create_mlflow_logger:
                self.loggers.append(MLflowLogger(self.log_dir, **(c.mlflow_logger_kwargs or {})))
        self.t_created = time.time()                   # reference point for recovery-time instrumentation
        if c.seconds_to_sleep and rank != 0:
            time.sleep(float(c.seconds_to_sleep))
        self.timing = TimingCallback() if c.log_step_timing else None
        self.checkpoint_callback = None
        if c.create_checkpoint_callback:
            cp = c.checkpoint_callback_params
            dirpath = cp.dirpath or os.path.join(self.log_dir, "checkpoints")
            self.checkpoint_callback = ModelCheckpoint(dirpath, cp, c.save_xser, c.async_checkpointing)
        trainer.exp_manager = self
        trainer.log_dir = self.log_dir

    def _setup_text_logging(self, rank: int, local_rank: int):
        c = self.cfg
        root = logging.getLogger("nxdt")
        root.setLevel(logging.INFO)
        if not root.handlers:
            h = logging.StreamHandler(sys.stdout)
            h.setFormatter(logging.Formatter("[%(asctime)s rank" + str(rank) + " %(levelname)s] %(message)s"))
            if rank == 0:
                root.addHandler(h)
        write_file = not ((c.log_local_rank_0_only and local_rank != 0) or (c.log_global_rank_0_only and rank != 0))
        if write_file:
            fh = logging.FileHandler(os.path.join(self.log_dir, f"nemo_log_globalrank-{rank}_localrank-{local_rank}.txt"))
            fh.setFormatter(logging.Formatter("[%(asctime)s %(levelname)s] %(message)s"))
            root.addHandler(fh)
            self._file_handler = fh
        if rank == 0:
            eh = logging.FileHandler(os.path.join(self.log_dir, "nemo_error_log.txt"))
            eh.setLevel(logging.WARNING)
            root.addHandler(eh)

    def log_metrics(self, metrics: Dict[str, float], step: int):
        for lg in self.loggers:
            lg.log_metrics(metrics, step)

    def close(self):
        for lg in self.loggers:
            lg.close()


def exp_manager(trainer, cfg_node: Optional[Config]) -> ExpManager:
    return ExpManager(trainer, cfg_node)
