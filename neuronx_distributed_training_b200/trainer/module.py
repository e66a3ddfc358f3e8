"""Task-module base class: the extension point users subclass to bring a new model.

Contract parity with the reference's ``BaseModelModule`` (reference
lightning_modules/model/base.py:49-712 and docs/developer_guides/new_model_guide.rst:40-113):
subclasses implement ``build_model()``, ``setup_optimizer_param_groups()``, ``init_weights(module,
device)`` and may override ``model_fwd_calc_loss``, ``forward_backward_step``, ``process_global_batch``.

One ``training_step`` = one optimizer step = one *global* batch: the per-DP-rank batch arrives on the
host, is copied H2D once, split into microbatches, each microbatch runs fwd+bwd with the loss scaled
by ``1/num_microbatches``, then the (ZeRO-1) optimizer steps.  Gradient accumulation is only ever
expressed through GBS/MBS (reference nlp_overrides.py:949-952).
"""
from __future__ import annotations

import contextlib
import logging
import os
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from .. import ops
from ..config import Config, ParallelDims, PrecisionPolicy
from ..data.base import get_batch_on_this_context_parallel_rank
from ..optim import DistributedOptimizer, get_optimizer, prepare_lr_scheduler
from ..parallel import state as ps
from ..parallel.layers import param_is_tp_duplicate
from ..parallel.random import set_random_seed
from ..utils.throughput import Throughput
from ..utils.profiling import nvtx_range

log = logging.getLogger("nxdt")


def get_param_groups_by_weight_decay(named_params, weight_decay: float = 0.01, no_decay_norm: bool = True):
    """bias (and optionally norm weights) -> no decay (reference utils/model_utils.py:4-22)."""
    decay, no_decay = [], []
    for n, p in named_params:
        if not p.requires_grad:
            continue
        is_norm = ("norm" in n.lower()) or ("LayerNorm" in n)
        if n.endswith("bias") or ".bias" in n or (no_decay_norm and is_norm):
            no_decay.append(p)
        else:
            decay.append(p)
    groups = [{"params": decay, "weight_decay": weight_decay}]
    if no_decay:
        groups.append({"params": no_decay, "weight_decay": 0.0})
    return groups


def keep_gathered_auto(tp: int, pp: int, tokens: int, column_input_widths, device_memory: int):
    """The ``keep_gathered_input: auto`` rule -> (keep?, extra bytes per micro-batch in flight).  Keeping the gathered input of
    every sequence-parallel column linear costs ``tokens * width * 2 * (tp - 1) / tp`` bytes each; it is switched on at TP >= 8
    (the measured case: +8.6 % tokens/s on Llama-3-8B) when that stays below 8 % of the device memory with ``pp`` micro-batches
    in flight (1F1B).  Llama-3-70B at TP=8 (19 GB on top of a 157 GB step) stays off."""
    extra = sum(tokens * int(k) * 2 * (tp - 1) // tp for k in column_input_widths)
    keep = tp >= 8 and extra > 0 and extra * max(1, pp) <= 0.08 * device_memory
    return keep, extra


class BaseModelModule:
    def __init__(self, cfg: Config, trainer=None):
        self.cfg = cfg
        self.config = cfg.model
        self.trainer = trainer
        ds = cfg.get("distributed_strategy", Config({}))
        self.dims = ParallelDims(world_size=ps.get_world_size(), tp=ps.get_tensor_model_parallel_size(),
                                 pp=ps.get_pipeline_model_parallel_size(), cp=ps.get_context_model_parallel_size(),
                                 ep=ps.get_expert_model_parallel_size(),
                                 vpp=int(ds.get("virtual_pipeline_model_parallel_size", 1) or 1))
        self.micro_batch_size = int(cfg.data.micro_batch_size)
        self.global_batch_size = int(cfg.data.global_batch_size)
        self.num_microbatches = self.global_batch_size // (self.micro_batch_size * self.dims.dp)
        self.sequence_parallel = bool(ds.get("sequence_parallel", False)) and self.dims.tp > 1
        self.zero1 = bool(ds.get("zero1", True))
        self.precision = PrecisionPolicy.from_config(cfg)
        self.seed = int(cfg.get("seed", 1234))
        set_random_seed(self.seed)
        em = cfg.get("exp_manager", Config({}))
        self.log_parameter_norm = bool(em.get("log_parameter_norm", False))
        self.log_gradient_norm = bool(em.get("log_gradient_norm", False))
        self.save_logits = bool(self.config.get("save_logits", False))
        self.model: Optional[nn.Module] = None
        self.optimizer: Optional[DistributedOptimizer] = None
        self.lr_scheduler = None
        self.throughput = Throughput(self.micro_batch_size, self.dims.dp, self.num_microbatches, 10)
        self.device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self._logged: Dict[str, float] = {}
        # CUDA-graph replay of the micro-batch forward+backward (utils/cuda_graph.py): opt-in
        fus = self.config.get("fusions", Config({})) if "fusions" in self.config else Config({})
        env = os.environ.get("NXDT_CUDA_GRAPH")
        self.use_cuda_graph = (bool(fus.get("cuda_graph", False)) if env is None else env == "1") \
            and self.device.type == "cuda" and self.dims.pp == 1
        self.cuda_graph_warmup_steps = int(os.environ.get("NXDT_CUDA_GRAPH_WARMUP", "1"))
        self._mb_graph = None
        self._eager_train_steps = 0
        # two-stream micro-batches (model.fusions.microbatch_overlap / NXDT_MB_OVERLAP): pairs of micro-batches are issued on
        # two CUDA streams so that the tensor-parallel collectives of one overlap the compute of the other.  The persistent
        # GEMM grid is capped (``comm_sms`` SMs stay free) so that a collective does not queue behind a resident GEMM.
        self.comm_sms = int(fus.get("comm_sms", os.environ.get("NXDT_COMM_SMS", 16)))
        env = os.environ.get("NXDT_MB_OVERLAP")
        self.use_mb_overlap = (bool(fus.get("microbatch_overlap", False)) if env is None else env == "1") \
            and self.device.type == "cuda" and self.dims.pp == 1 and not self.use_cuda_graph
        self._mb_streams = None

    # ---- to be provided by subclasses -------------------------------------------------------------
    def build_model(self) -> nn.Module:
        raise NotImplementedError

    def setup_optimizer_param_groups(self) -> List[Dict[str, Any]]:
        wd = float(self.config.optim.get("weight_decay", 0.01))
        return get_param_groups_by_weight_decay(self.local_named_parameters(), wd)

    def init_weights(self, module: nn.Module, device) -> None:
        for fn in ("init_weight_cpu", "initialize_weight_biases", "reset_parameters"):
            if hasattr(module, fn):
                getattr(module, fn)()
                return

    # ---- lifecycle ---------------------------------------------------------------------------------
    def setup(self, stage: str = "fit"):
        fus = self.config.get("fusions", Config({})) if "fusions" in self.config else Config({})
        kernels = str(self.cfg.get("kernels", "fused") or "fused").lower()
        if kernels not in ("fused", "baseline"):
            raise ValueError(f"kernels: expected 'fused' or 'baseline', got {kernels!r}")
        if kernels == "baseline":
            # top-level `kernels: baseline`: every hot op on the library path (NCCL + cuBLAS + SDPA + unfused torch ops) --
            # the A/B arm against which the sm_100a kernels are measured (same as NXDT_OPS_BACKEND=torch)
            ops.set_backend("torch")
        if "fp8_gemm" in fus or str(self.cfg.get("precision", Config({})).get("gemm_dtype", "")).lower() in ("fp8", "e4m3"):
            # linear-layer GEMMs in e4m3 with row / column block scales (ops/fp8.py); LM head, attention, norms stay bf16
            ops.fp8.set_enabled(bool(fus.get("fp8_gemm", True)))
            if "fp8_recipe" in fus:            # "mx" (default: 1x32 E8M0 blocks, scaled in the tensor core) | "row" (1xK fp32 scales)
                ops.fp8.set_recipe(str(fus.get("fp8_recipe")))
        if self.model is None:
            self.model = self.build_model()
            self._maybe_enable_fused_tp()
        self._configure_keep_gathered(fus)
        self._maybe_enable_nvls(fus)

    def _maybe_enable_nvls(self, fus) -> int:
        """``model.fusions.nvls_norm_allgather`` (default false; ``NXDT_NVLS=1``): the sequence-parallel RMSNorm in front of
        a column linear and the all-gather that follows run as ONE kernel that multicasts the normalised rows through the
        NVSwitch (``multimem.st``, parallel/nvls.py) -- for the linears the fused all-gather -> GEMM kernel leaves on NCCL."""
        want = bool(fus.get("nvls_norm_allgather", False)) if "nvls_norm_allgather" in fus else False
        if "NXDT_NVLS" in os.environ:
            want = os.environ["NXDT_NVLS"] == "1"
        self.nvls_layers = 0
        if not want or self.device.type != "cuda" or self.dims.tp < 2 or not self.sequence_parallel or self.model is None:
            return 0
        if self.use_cuda_graph or self.use_mb_overlap:
            # the arrival targets are host-side counters (a replayed graph would repeat them) and the workspace is single-stream
            log.info("nvls_norm_allgather: off with cuda_graph / microbatch_overlap")
            return 0
        from ..parallel.nvls import enable_nvls_norm_allgather
        seq = int(self.config.get("encoder_seq_length", None) or self.cfg.data.get("seq_length", 8192))
        self.nvls_layers = enable_nvls_norm_allgather(self.model, max(1, seq // self.dims.cp), self.micro_batch_size)
        if ps.get_rank() == 0:
            log.info("NVLS norm + all-gather attached to %d linears", self.nvls_layers)
        return self.nvls_layers

    def _configure_keep_gathered(self, fus) -> None:
        """``model.fusions.keep_gathered_input``: true | false | auto (default).  Column linears of a sequence-parallel
        layer either keep their all-gathered input for the weight gradient or gather it a second time in backward.
        Measured on 8 x B200, Llama-3-8B, TP=8: keeping it is +8.6 % tokens/s for +3.5 GB (the layer is transfer-bound
        at large TP).  ``auto`` keeps it when TP >= 8 (the measured case) and the extra activations stay below 8 % of the device memory
        (Llama-3-70B at TP=8 would add 19 GB to a 157 GB step: stays off)."""
        from ..parallel.layers import ColumnParallelLinear, GQAQKVColumnParallelLinear, set_keep_gathered_input
        want = fus.get("keep_gathered_input", "auto") if "keep_gathered_input" in fus else "auto"
        if "NXDT_KEEP_GATHERED" in os.environ:
            want = os.environ["NXDT_KEEP_GATHERED"] == "1"
        if isinstance(want, str):
            want = {"true": True, "false": False}.get(want.lower(), "auto")
        if want == "auto":
            want = False
            if self.device.type == "cuda" and self.sequence_parallel and self.model is not None:
                seq = int(self.config.get("encoder_seq_length", None) or self.cfg.data.get("seq_length", 8192))
                tokens = self.micro_batch_size * max(1, seq // self.dims.cp)
                widths = [int(getattr(m, "input_size", None) or m.weight.shape[-1]) for m in self.model.modules()
                          if isinstance(m, (ColumnParallelLinear, GQAQKVColumnParallelLinear))
                          and getattr(m, "sequence_parallel_enabled", False)]
                total = torch.cuda.get_device_properties(self.device).total_memory
                want, extra = keep_gathered_auto(self.dims.tp, self.dims.pp, tokens, widths, total)
                if ps.get_rank() == 0:
                    log.info("keep_gathered_input=auto -> %s (%.1f GB of gathered inputs per micro-batch in flight)",
                             want, extra / 1e9)
        set_keep_gathered_input(bool(want))

    def _maybe_enable_fused_tp(self) -> int:
        """Attach the fused AG->GEMM / GEMM->RS / AG-wgrad kernels (parallel/fused.py) to the sequence-parallel linears.
        ``model.fusions.fused_tp``: ``auto`` (default: on for CUDA runs with TP > 1 and sequence parallelism, quietly
        off when the workspace cannot be set up), ``true`` (errors are fatal), ``false``.  ``NXDT_FUSED_TP=0|1`` overrides."""
        fus = self.config.get("fusions", Config({})) if "fusions" in self.config else Config({})
        want = fus.get("fused_tp", "auto")
        env = os.environ.get("NXDT_FUSED_TP")
        if env is not None:
            want = env != "0"
        if isinstance(want, str):
            want = {"true": True, "false": False}.get(want.lower(), "auto")
        self.fused_tp_layers = 0
        if want is False or self.device.type != "cuda" or self.dims.tp < 2 or not self.sequence_parallel:
            return 0
        if ops.get_backend() == "torch" or self.use_mb_overlap:      # library baseline / two-stream mode: NCCL TP path
            return 0
        if ops.fp8.enabled() and want == "auto":
            # fp8 linears: e4m3 GEMMs + e4m3 all-gathers over NCCL (the fused kernels are bf16)
            log.info("fp8_gemm: tensor-parallel collectives stay on NCCL (activations are gathered as e4m3 bytes)")
            return 0
        if self.use_cuda_graph and want == "auto" and os.environ.get("NXDT_CUDA_GRAPH_WITH_FUSED_TP", "0") != "1":
            log.info("cuda_graph requested: tensor-parallel collectives stay on NCCL (graph replay of the fused kernels is "
                     "not validated yet, DESIGN.md 3.2b)")
            return 0
        try:
            from ..parallel.fused import enable_fused_tp
            seq = int(self.config.get("encoder_seq_length", None) or self.cfg.data.get("seq_length", 8192))
            self.fused_tp_layers = enable_fused_tp(self.model, seq_len=max(1, seq // self.dims.cp),
                                                   micro_batch=self.micro_batch_size)
            if ps.get_rank() == 0:
                log.info("fused tensor-parallel kernels attached to %d linears", self.fused_tp_layers)
        except Exception as e:  # noqa: BLE001 - e.g. no peer access between the GPUs of the TP group
            if want is True:
                raise
            log.warning("fused tensor-parallel kernels not enabled (%s); using the NCCL path", e)
        return self.fused_tp_layers

    def local_named_parameters(self):
        m = self.model
        if hasattr(m, "local_named_parameters"):
            return list(m.local_named_parameters())
        return list(m.named_parameters())

    def configure_optimizers(self, max_steps: Optional[int] = None):
        ocfg = self.config.optim
        algo, extra = get_optimizer(str(ocfg.get("name", "adamw")))
        groups = self.setup_optimizer_param_groups()
        clip = self.cfg.trainer.get("gradient_clip_val", None)
        pol = self.precision
        eps = ocfg.get("eps", 1e-8)
        eps = tuple(float(e) for e in eps) if isinstance(eps, (list, tuple)) else float(eps)     # adafactor: (eps1, eps2)
        known = {"name", "lr", "betas", "eps", "weight_decay", "momentum", "sched", "overlap_grad_reduce", "capturable"}
        extras = {k: ocfg[k] for k in ocfg.keys() if k not in known} if hasattr(ocfg, "keys") else {}
        common = dict(lr=float(ocfg.get("lr", 1e-4)), betas=tuple(ocfg.get("betas", (0.9, 0.999))),
                      eps=eps, weight_decay=float(ocfg.get("weight_decay", 0.01)), extra_defaults=extras,
                      algorithm=algo, momentum=float(ocfg.get("momentum", 0.0)), zero1=self.zero1,
                      master_weights=pol.master_weights, fp32_grad_acc=pol.fp32_grad_acc, max_grad_norm=clip,
                      stochastic_rounding=pol.stochastic_rounding, state_dtype=pol.optimizer_state_dtype)
        is_expert = lambda p: bool(getattr(p, "expert_parallel", False))
        if self.dims.ep > 1 and any(is_expert(p) for g in groups for p in g["params"]):
            # expert-parallel parameters differ across the EP group: reduce them over the expert-data-parallel group
            # only, and add their norm over the EP group; everything else reduces over the full DP(xCP) group.
            from ..optim.distributed import ChainedOptimizer
            dense = [dict(g, params=[p for p in g["params"] if not is_expert(p)]) for g in groups]
            expert = [dict(g, params=[p for p in g["params"] if is_expert(p)]) for g in groups]
            dense = [g for g in dense if g["params"]]
            expert = [g for g in expert if g["params"]]
            o_dense = DistributedOptimizer(dense, **common)
            # tokens of every EP rank reach an expert, each carrying d(loss_rank): scale by 1/EP so that together with the
            # mean over the expert-data-parallel group the gradient is the mean over the whole data-parallel batch
            o_exp = DistributedOptimizer(expert, dp_group=ps.get_expert_data_parallel_group() or "self",
                                         extra_grad_scale=1.0 / self.dims.ep, **common)
            self.optimizer = ChainedOptimizer([o_dense, o_exp], [(), (ps.get_expert_model_parallel_group(),)])
        else:
            self.optimizer = DistributedOptimizer(groups, **common)
        # overlap the data-parallel gradient reduction with the last micro-batch's backward (non-pipelined path);
        # bucket size = the reference's `bucket_size_collectives` (MB)
        if bool(ocfg.get("overlap_grad_reduce", False)) or os.environ.get("NXDT_OVERLAP_GRAD_REDUCE", "0") == "1":
            if self.dims.pp == 1 and isinstance(self.optimizer, DistributedOptimizer):
                if self.optimizer.enable_grad_reduce_overlap(float(self.cfg.get("bucket_size_collectives", 256) or 256)):
                    log.info("data-parallel gradient reduction overlaps backward (%d buckets)", len(self.optimizer._buckets))
        self._maybe_enable_fused_zero1(ocfg)
        sched = ocfg.get("sched", None)
        self.lr_scheduler = prepare_lr_scheduler(self.optimizer, sched, max_steps) if sched is not None else None
        return self.optimizer, self.lr_scheduler

    def _maybe_enable_fused_zero1(self, ocfg) -> bool:
        """ZeRO-1 step as two peer-memory kernels (gradient reduce-scatter by peer pulls -> AdamW on the shard -> parameter
        all-gather by peer pushes, ``optim/fused_zero1.py``) instead of reduce_scatter + per-segment Adam + all_gather.
        ``model.optim.fused_zero1: auto | true | false`` (``NXDT_FUSED_ZERO1=0|1``); needs AdamW with fp32 master weights and
        state, a data-parallel group of 2..8 GPUs with peer access (one node), and no overlapped bucket reduction."""
        want = ocfg.get("fused_zero1", "auto")
        env = os.environ.get("NXDT_FUSED_ZERO1")
        if env is not None:
            want = env != "0"
        if isinstance(want, str):
            want = {"true": True, "false": False}.get(want.lower(), "auto")
        self.fused_zero1 = False
        opt = self.optimizer
        if want is False or self.device.type != "cuda" or not isinstance(opt, DistributedOptimizer) or ops.get_backend() == "torch":
            return False
        if not opt.zero1 or not (2 <= opt.dp_size <= 8) or getattr(opt, "_buckets", None):
            return False
        if opt.master is None or opt.algorithm != "adamw" or opt.exp_avg.dtype != torch.float32:
            return False
        if any(getattr(p, "norm_duplicate_range", None) is not None for p in opt.params):
            return False            # KV-head replicas: the fused kernels' norm mask has no per-row exclusion
        try:
            from ..optim.fused_zero1 import enable_fused_zero1
            self.fused_zero1 = bool(enable_fused_zero1(opt))
            if self.fused_zero1 and ps.get_rank() == 0:
                log.info("ZeRO-1 step runs as fused peer-memory kernels over the data-parallel group (%d GPUs)", opt.dp_size)
        except Exception as e:  # noqa: BLE001 - e.g. the DP group spans nodes (no peer mapping)
            if want is True:
                raise
            log.warning("fused ZeRO-1 step not enabled (%s); using NCCL reduce-scatter / all-gather", e)
        return self.fused_zero1

    # ---- step ----------------------------------------------------------------------------------------
    def autocast_ctx(self):
        if self.precision.autocast:
            return torch.autocast(device_type=self.device.type, dtype=torch.bfloat16)
        return contextlib.nullcontext()

    def to_device(self, batch: Dict[str, Any]) -> Dict[str, Any]:
        return {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}

    def get_batch_iterator(self, batch: Dict[str, torch.Tensor]):
        mbs = self.micro_batch_size
        n = self.num_microbatches
        for k in range(n):
            yield {key: (v[k * mbs:(k + 1) * mbs] if torch.is_tensor(v) else v) for key, v in batch.items()}

    def model_fwd_calc_loss(self, batch: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, Any]:
        extra = {"return_logits": True} if self.save_logits and self.dims.pp == 1 else {}
        out = self.model(input_ids=batch["input_ids"], labels=batch["labels"],
                         attention_mask=batch.get("attention_mask"), loss_mask=batch.get("loss_mask"), **extra)
        if self.save_logits and out[1] is not None:
            # model.save_logits (reference base.py:68,381): the (vocab-parallel) logits of the latest micro-batch stay
            # available for numerics comparisons; dumped next to the logs when NXDT_DUMP_LOGITS_DIR is set
            self.last_logits = out[1].detach()
            dump = os.environ.get("NXDT_DUMP_LOGITS_DIR")
            if dump:
                os.makedirs(dump, exist_ok=True)
                torch.save(self.last_logits.cpu(), os.path.join(dump, f"logits_rank{ps.get_rank()}.pt"))
        return out[0], out[1]

    def forward_backward_step(self, batch: Dict[str, torch.Tensor], is_training: bool = True) -> torch.Tensor:
        if self.dims.pp > 1:
            fn = self.model.run_train if is_training else self.model.run_eval
            running = fn(**batch)
        else:
            running = torch.zeros((), dtype=torch.float32, device=self.device)
            if is_training and self.use_mb_overlap and self.num_microbatches >= 2:
                running = self._overlapped_microbatches(batch)
                batch = {}          # nothing left for the sequential loop below
            n_seen = 0
            for mb in (self.get_batch_iterator(batch) if batch else ()):
                n_seen += 1
                if is_training and n_seen == self.num_microbatches and self.optimizer is not None \
                        and hasattr(self.optimizer, "start_grad_sync_window"):
                    self.optimizer.start_grad_sync_window()      # DP reduction of finished buckets overlaps this backward
                if is_training and self._graphed_microbatch(mb, running):
                    continue
                with contextlib.nullcontext() if is_training else torch.no_grad():
                    with nvtx_range("forward"):
                        loss, _ = self.model_fwd_calc_loss(mb)
                    loss = loss.float() / self.num_microbatches
                if is_training:
                    with nvtx_range("backward"):
                        loss.backward()
                running += loss.detach()
            if is_training:
                self._eager_train_steps += 1
        # loss mean over CP then DP (reference base.py:387-395)
        cpg, dpg = ps.get_context_model_parallel_group(), ps.get_data_parallel_group()
        if ps.group_size(cpg) > 1:
            dist.all_reduce(running, group=cpg)
            running /= ps.group_size(cpg)
        if ps.group_size(dpg) > 1:
            dist.all_reduce(running, group=dpg)
            running /= ps.group_size(dpg)
        return running

    def _overlapped_microbatches(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Forward + backward of the micro-batches, two at a time on two side streams.  Both streams share the flat
        ``main_grad`` buffer: the wgrad GEMMs add into it with element-wise atomic TMA reductions, the
        read-modify-write accumulations are ordered by ``parallel.grad_hooks.ordered_main_grad_update``; per-stream scratch in the
        fused LM-head / cross-entropy keeps the two micro-batches apart.  Requires NCCL tensor parallelism (the fused
        TP kernels own every SM and one symmetric workspace, so they cannot run two at a time)."""
        from ..parallel import fused as fused_mod
        if fused_mod._FTP_CACHE:
            raise RuntimeError("microbatch_overlap needs the NCCL tensor-parallel path (do not call enable_fused_tp)")
        from ..parallel.grad_hooks import set_multi_stream
        set_multi_stream(True)
        main = torch.cuda.current_stream()
        if self._mb_streams is None:
            self._mb_streams = [torch.cuda.Stream(), torch.cuda.Stream()]
            C = ops.C()
            if C is not None and hasattr(C, "gemm_set_cta_limit") and self.comm_sms > 0:
                sms = torch.cuda.get_device_properties(self.device).multi_processor_count
                C.gemm_set_cta_limit(max(1, sms - self.comm_sms))
        streams = self._mb_streams
        partial = [torch.zeros((), dtype=torch.float32, device=self.device) for _ in streams]
        for s in streams:
            s.wait_stream(main)
        mbs = list(self.get_batch_iterator(batch))
        for k in range(0, len(mbs), 2):
            losses = []
            for idx, mb in enumerate(mbs[k:k + 2]):
                with torch.cuda.stream(streams[idx]):
                    with nvtx_range("forward"):
                        loss, _ = self.model_fwd_calc_loss(mb)
                    losses.append(loss.float() / self.num_microbatches)
            for idx, loss in enumerate(losses):
                with torch.cuda.stream(streams[idx]):
                    with nvtx_range("backward"):
                        loss.backward()
                    partial[idx] += loss.detach()
        for s in streams:
            main.wait_stream(s)
        return partial[0] + partial[1]

    def _graphed_microbatch(self, mb: Dict[str, Any], running: torch.Tensor) -> bool:
        """Replay (capturing first if needed) the CUDA graph of one micro-batch; False -> run it eagerly."""
        if not self.use_cuda_graph or self._eager_train_steps < self.cuda_graph_warmup_steps:
            return False
        from ..utils.cuda_graph import MicrobatchGraph
        if self._mb_graph is None:
            def fwd_bwd(static_mb):
                loss, _ = self.model_fwd_calc_loss(static_mb)
                loss = loss.float() / self.num_microbatches
                loss.backward()
                return loss
            g = MicrobatchGraph(fwd_bwd)
            try:
                g.capture(mb)
            except Exception as e:  # noqa
                log.warning("CUDA-graph capture of the micro-batch failed (%s); continuing eagerly", e,
                            exc_info=os.environ.get("NXDT_CUDA_GRAPH_DEBUG", "0") == "1")
                self.use_cuda_graph = False
                return False
            self._mb_graph = g
            if ps.get_rank() == 0:
                log.info("captured the micro-batch forward+backward as a CUDA graph")
        if not self._mb_graph.matches(mb):
            return False
        running += self._mb_graph.run(mb)
        return True

    def training_step(self, batch: Dict[str, torch.Tensor], batch_idx: int = 0) -> torch.Tensor:
        self.optimizer.zero_grad()
        seq = self.config.get("encoder_seq_length", None)
        if seq is not None and "input_ids" in batch:
            assert batch["input_ids"].shape[1] == int(seq), \
                f"batch sequence length {batch['input_ids'].shape[1]} != model.encoder_seq_length {seq}"
        batch = self.to_device(batch)
        batch = get_batch_on_this_context_parallel_rank(batch)
        with self.autocast_ctx():
            loss_mean = self.forward_backward_step(batch, is_training=True)
        ppg = ps.get_pipeline_model_parallel_group()
        if ps.group_size(ppg) > 1:
            dist.all_reduce(loss_mean, group=ppg)       # last stage holds the loss, the others 0
        self.sync_special_grads()
        with nvtx_range("optimizer_step"):
            self.optimizer.step()
        if self.lr_scheduler is not None:
            self.lr_scheduler.step()
        return loss_mean

    def sync_special_grads(self):
        """KV-replica gradient sums (GQA with kv_replicator > 1)."""
        for m in self.model.modules():
            if hasattr(m, "sync_kv_replica_grads"):
                m.sync_kv_replica_grads()

    @torch.no_grad()
    def validation_step(self, batch: Dict[str, torch.Tensor], batch_idx: int = 0) -> torch.Tensor:
        batch = get_batch_on_this_context_parallel_rank(self.to_device(batch))
        self.model.eval()
        try:
            with self.autocast_ctx():
                loss = self.forward_backward_step(batch, is_training=False)
        finally:
            self.model.train()
        ppg = ps.get_pipeline_model_parallel_group()
        if ps.group_size(ppg) > 1:
            dist.all_reduce(loss, group=ppg)
        return loss

    @torch.no_grad()
    def predict_step(self, batch: Dict[str, torch.Tensor], batch_idx: int = 0) -> Dict[str, torch.Tensor]:
        """Greedy next-token predictions and per-token log-probs of the labels for one batch (non-pipelined models)."""
        if self.dims.pp > 1:
            raise NotImplementedError("predict_step is implemented for pipeline_model_parallel_size == 1")
        from ..parallel.loss import from_parallel_logits_to_logprobs
        batch = get_batch_on_this_context_parallel_rank(self.to_device(batch))
        self.model.eval()
        try:
            with self.autocast_ctx():
                _, logits = self.model(input_ids=batch["input_ids"], labels=None)          # [B, S, V/TP]
        finally:
            self.model.train()
        tp_group = ps.get_tensor_model_parallel_group()
        local_max, local_arg = logits.float().max(dim=-1)
        vocab_start = ps.get_tensor_model_parallel_rank() * logits.shape[-1]
        if ps.group_size(tp_group) > 1:
            gmax = local_max.clone()
            dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=tp_group)
            cand = torch.where(local_max == gmax, local_arg + vocab_start, torch.full_like(local_arg, 1 << 40))
            dist.all_reduce(cand, op=dist.ReduceOp.MIN, group=tp_group)
            pred = cand
        else:
            pred = local_arg
        out = {"predictions": pred}
        if "labels" in batch and batch["labels"] is not None:
            labels = batch["labels"]
            shifted = torch.cat([labels[:, 1:], labels.new_full((labels.shape[0], 1), -100)], dim=1)
            out["logprobs"] = from_parallel_logits_to_logprobs(logits, shifted)
        return out

    # ---- metrics -------------------------------------------------------------------------------------
    @torch.no_grad()
    def parameter_norm(self, norm_type: float = 2.0) -> torch.Tensor:
        """p-norm (``math.inf`` for the max norm) of the model parameters, every logical element counted once:
        tensor-parallel duplicates and KV-head replicas are filtered, then the partial results are reduced over the
        TP, CP and PP groups in that order, as the reference does (base.py:397-452).  Parameters are REPLICATED over
        the context-parallel group, so the CP reduction is a MAX for the inf-norm and SUM / cp for p-norms (the
        reference's plain SUM over CP counts every element cp times)."""
        import math
        tp_rank = ps.get_tensor_model_parallel_rank()
        norm_type = float(norm_type)
        total = torch.zeros((), dtype=torch.float32, device=self.device)
        for _, p in self.local_named_parameters():
            if param_is_tp_duplicate(p) and tp_rank != 0:
                continue
            x = p.detach().float()
            if norm_type == math.inf:
                total = torch.maximum(total, x.abs().max())
                continue
            total += x.pow(2).sum() if norm_type == 2.0 else x.abs().pow(norm_type).sum()
            dup = getattr(p, "norm_duplicate_range", None)          # KV-head replica rows another rank already counts
            if dup is not None:
                d = x.reshape(-1)[dup[0]:dup[1]]
                total -= d.pow(2).sum() if norm_type == 2.0 else d.abs().pow(norm_type).sum()
        op = dist.ReduceOp.MAX if norm_type == math.inf else dist.ReduceOp.SUM
        for kind, grp in (("tp", ps.get_tensor_model_parallel_group()), ("cp", ps.get_context_model_parallel_group()),
                          ("pp", ps.get_pipeline_model_parallel_group())):
            if ps.group_size(grp) > 1:
                dist.all_reduce(total, op=op, group=grp)
                if kind == "cp" and norm_type != math.inf:
                    total /= ps.group_size(grp)
        if norm_type == math.inf:
            return total
        return total.sqrt() if norm_type == 2.0 else total.pow(1.0 / norm_type)

    def collect_metrics(self, loss: torch.Tensor, global_step: int, consumed_samples: int) -> Dict[str, float]:
        m = {"reduced_train_loss": float(loss), "global_step": float(global_step),
             "consumed_samples": float(consumed_samples)}
        if self.optimizer is not None:
            m["lr"] = float(self.optimizer.param_groups[0]["lr"])
            if self.log_gradient_norm and self.optimizer.grad_norm is not None:
                m["gradient_norm"] = float(self.optimizer.grad_norm)
        if self.log_parameter_norm:
            m["parameter_norm"] = float(self.parameter_norm())
        m.update({k: (float(v) if torch.is_tensor(v) else v) for k, v in self._logged.items()})
        m["throughput"] = self.throughput.get_throughput()
        m["throughput_peak"] = self.throughput.throughput_peak
        return m

    # ---- hooks the trainer calls -----------------------------------------------------------------------
    def on_train_start(self, trainer):
        pass

    def on_train_end(self, trainer):
        pass

    def extra_checkpoint_files(self) -> Dict[str, str]:
        """Small text files stored in the checkpoint directory next to ``user_content.pt``: with PEFT and
        ``save_lora_config_adapter`` the LoRA description as ``adapter_config.json``."""
        a = self.cfg.get("model_alignment_strategy", None)
        peft = a.get("peft", None) if a is not None else None
        if peft is None or not bool(peft.get("save_lora_config_adapter", True)):
            return {}
        import json
        from ..models.lora import LoraConfig
        base = self.config.get("model_config", None)
        return {"adapter_config.json": json.dumps(LoraConfig.from_config(peft).to_adapter_config(str(base) if base else None),
                                                  indent=2)}

    def state_dict(self) -> Dict[str, torch.Tensor]:
        a = self.cfg.get("model_alignment_strategy", None)
        peft = a.get("peft", None) if a is not None else None
        if peft is not None:
            from ..models import lora
            if bool(peft.get("merge_lora", False)):
                return {k: v for c in self.model.chunks for k, v in lora.merged_state_dict(c).items()}
            if not bool(peft.get("save_lora_base", False)):
                return {k: v for c in self.model.chunks for k, v in lora.lora_state_dict(c).items()}
        return self.model.state_dict()

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = False):
        """Non-strict load that ignores rotary ``inv_freq`` buffers but raises on any other missing key
        unless PEFT is active (reference nlp_overrides.py:1157-1200)."""
        peft = self.cfg.get("model_alignment_strategy", None) is not None and \
            self.cfg.model_alignment_strategy.get("peft", None) is not None
        if peft:
            # checkpoints of the base model use the un-wrapped names: route them to ``<module>.base_layer.*``
            own = self.model.state_dict()
            remapped = dict(sd)
            for k in own:
                if ".base_layer." in k:
                    plain = k.replace(".base_layer.", ".")
                    if plain in sd and k not in sd:
                        remapped[k] = remapped.pop(plain)
            sd = remapped
        res = self.model.load_state_dict(sd, strict=False)
        missing = [k for k in res.missing_keys if not k.endswith("rotary_emb.inv_freq")]
        if missing and not peft:
            raise RuntimeError(f"missing keys when loading checkpoint: {missing[:8]}{' ...' if len(missing) > 8 else ''}")
        return res
