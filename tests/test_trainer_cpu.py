import glob
import json
import os

import pytest
import torch

from neuronx_distributed_training_b200.config import from_dict, process_config
from neuronx_distributed_training_b200.parallel import state as ps


def _cfg(log_dir, max_steps=4, **model_kw):
    model = {"model_config": "llama_125m", "encoder_seq_length": 32, "num_layers": 2, "hidden_size": 64, "intermediate_size": 128,
             "num_attention_heads": 4, "num_kv_heads": 2, "vocab_size": 128, "fusions": {"flash_attention": False},
             "optim": {"name": "adamw_fp32OptState", "lr": 1e-2, "weight_decay": 0.01, "betas": [0.9, 0.95],
                       "sched": {"name": "LinearAnnealingWithWarmUp", "warmup_steps": 1, "max_steps": 10}}}
    model.update(model_kw)
    return from_dict({
        "name": "t", "model_source": "hf", "seed": 3,
        "trainer": {"devices": 1, "num_nodes": 1, "max_steps": max_steps, "log_every_n_steps": 1, "val_check_interval": 2,
                    "limit_val_batches": 2, "gradient_clip_val": 1.0},
        "exp_manager": {"explicit_log_dir": log_dir, "create_tensorboard_logger": False, "create_checkpoint_callback": True,
                        "resume_if_exists": True, "resume_ignore_no_checkpoint": True, "log_parameter_norm": True, "log_gradient_norm": True,
                        "checkpoint_callback_params": {"save_top_k": 1, "every_n_train_steps": 2, "filename": "t--{step}-{consumed_samples}"}},
        "distributed_strategy": {"tensor_model_parallel_size": 1, "zero1": True},
        "data": {"micro_batch_size": 1, "global_batch_size": 2, "synthetic": True},
        "model": model, "precision": {"type": "fp32"},
    })


@pytest.fixture(autouse=True)
def _mp():
    ps.initialize_model_parallel(1, 1, 1)
    yield
    ps.destroy_model_parallel()
    import logging
    lg = logging.getLogger("nxdt")
    for h in list(lg.handlers):
        lg.removeHandler(h)


def _fit(cfg):
    from neuronx_distributed_training_b200.trainer import Trainer
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    from neuronx_distributed_training_b200.data import SyntheticDataModule
    from neuronx_distributed_training_b200.utils.exp_manager import exp_manager
    process_config(cfg, 1)
    tr = Trainer(**cfg.trainer.to_dict())
    em = exp_manager(tr, cfg.exp_manager)
    mod = HFLlamaModule(cfg, tr)
    tr.fit(mod, SyntheticDataModule(cfg, tr), ckpt_path=em.resume_from_checkpoint)
    return tr, mod


def test_fit_logs_validates_and_checkpoints(tmp_path):
    tr, mod = _fit(_cfg(str(tmp_path / "run")))
    assert tr.global_step == 4
    m = tr.logged_metrics
    for k in ("reduced_train_loss", "lr", "gradient_norm", "parameter_norm", "consumed_samples", "throughput", "throughput_peak"):
        assert k in m, k
    assert m["consumed_samples"] == 8 and "val_loss" in tr.callback_metrics
    files = os.listdir(tmp_path / "run")
    assert "cmd-args.log" in files and "metrics.jsonl" in files and any(f.startswith("nemo_log_globalrank-0") for f in files)
    ck = glob.glob(str(tmp_path / "run" / "checkpoints" / "*.ckpt"))
    assert len(ck) == 1 and "step=4-consumed_samples=8.0" in ck[0]          # save_top_k=1 kept the newest
    lines = [json.loads(l) for l in open(tmp_path / "run" / "metrics.jsonl")]
    assert any("val_loss" in l for l in lines) and sum("reduced_train_loss" in l for l in lines) == 4


def test_trainer_rejects_unsupported_options():
    from neuronx_distributed_training_b200.trainer import Trainer
    with pytest.raises(ValueError):
        Trainer(accumulate_grad_batches=2)
    with pytest.raises(ValueError):
        Trainer(gradient_clip_algorithm="value")
    t = Trainer(max_time="00:00:01:30", lnc=2, sequential_move_factor=11)    # Neuron-only knobs are accepted
    assert t.max_time == 90.0


@pytest.mark.parametrize("gran,rec", [("selective", ["LlamaMLP", "CoreAttention"]), ("full", None)])
def test_activation_checkpointing_is_exact(tmp_path, gran, rec):
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    losses = {}
    for tag, kw in (("none", {}), ("ckpt", {"activations_checkpoint_granularity": gran, "activations_checkpoint_recompute": rec})):
        cfg = process_config(_cfg(str(tmp_path / tag), **kw), 1)
        mod = HFLlamaModule(cfg, None); mod.setup(); mod.configure_optimizers(10)
        ids = torch.randint(0, 128, (2, 32), generator=torch.Generator().manual_seed(0))
        ls = [float(mod.training_step({"input_ids": ids, "labels": ids.clone()})) for _ in range(3)]
        losses[tag] = ls
        if tag == "ckpt":
            n = sum(1 for m in mod.model.modules() if getattr(m, "_activation_checkpointed", False))
            assert n == (4 if gran == "selective" else 2)
    assert losses["none"] == pytest.approx(losses["ckpt"], abs=1e-6)


def test_swiglu_recompute_runs_inside_down_proj_node(tmp_path):
    """Selecting ActivationMultiplyMLP keeps only gate||up: no checkpoint wrapper, identical losses."""
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    losses = {}
    for tag, kw in (("none", {}), ("swiglu", {"activations_checkpoint_granularity": "selective",
                                              "activations_checkpoint_recompute": ["ActivationMultiplyMLP"]})):
        cfg = process_config(_cfg(str(tmp_path / tag), **kw), 1)
        mod = HFLlamaModule(cfg, None); mod.setup(); mod.configure_optimizers(10)
        ids = torch.randint(0, 128, (2, 32), generator=torch.Generator().manual_seed(0))
        losses[tag] = [float(mod.training_step({"input_ids": ids, "labels": ids.clone()})) for _ in range(3)]
        if tag == "swiglu":
            acts = [m for m in mod.model.modules() if type(m).__name__ == "ActivationMultiplyMLP"]
            assert len(acts) == 2 and all(m.recompute_in_consumer for m in acts)
            assert not any(getattr(m, "_activation_checkpointed", False) for m in mod.model.modules())
    assert losses["none"] == pytest.approx(losses["swiglu"], abs=1e-6)


def test_unknown_recompute_class_and_exp_manager_keys(tmp_path):
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    from neuronx_distributed_training_b200.utils.exp_manager import ExpManagerConfig
    cfg = process_config(_cfg(str(tmp_path / "x"), activations_checkpoint_granularity="selective", activations_checkpoint_recompute=["Nope"]), 1)
    with pytest.raises(ValueError):
        HFLlamaModule(cfg, None).setup()
    with pytest.raises(ValueError):
        ExpManagerConfig.from_config(from_dict({"not_a_key": 1}))


def test_exp_manager_small_options(tmp_path):
    """files_to_copy, save_bf16, enable_recovery_time_instrumentation, max_time_per_run."""
    src = tmp_path / "launch_note.txt"
    src.write_text("hello")
    cfg = _cfg(str(tmp_path / "run"), max_steps=2)
    cfg.exp_manager["files_to_copy"] = [str(src)]
    cfg.exp_manager["save_bf16"] = True
    cfg.exp_manager["enable_recovery_time_instrumentation"] = True
    cfg.exp_manager["max_time_per_run"] = "00:01:00:00"
    tr, mod = _fit(cfg)
    assert (tmp_path / "run" / "launch_note.txt").read_text() == "hello"
    assert tr.max_time == 3600.0 and tr.callback_metrics["recovery_time_s"] > 0
    from neuronx_distributed_training_b200.checkpoint import io as ckpt_io
    ck = sorted(glob.glob(str(tmp_path / "run" / "checkpoints" / "*.ckpt")))[-1]
    m, _, _ = ckpt_io.load_checkpoint(os.path.dirname(ck), os.path.basename(ck), optimizer=False)
    assert all(v.dtype == torch.bfloat16 for v in m.values() if torch.is_tensor(v) and v.is_floating_point())


def test_trainer_test_and_predict_loops(tmp_path):
    """Trainer.test / Trainer.predict (the reference's evaluation and prediction loops) on a trained checkpoint."""
    from neuronx_distributed_training_b200.trainer import Trainer
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    from neuronx_distributed_training_b200.data import SyntheticDataModule
    cfg = _cfg(str(tmp_path / "run"), max_steps=2)
    tr, mod = _fit(cfg)
    ck = sorted(glob.glob(str(tmp_path / "run" / "checkpoints" / "*.ckpt")))[-1]
    # a fresh trainer + module evaluates the saved weights
    cfg2 = process_config(_cfg(str(tmp_path / "eval"), max_steps=2), 1)
    cfg2.trainer["limit_test_batches"] = 2
    tr2 = Trainer(**cfg2.trainer.to_dict())
    mod2 = HFLlamaModule(cfg2, tr2)
    dm2 = SyntheticDataModule(cfg2, tr2)
    if not hasattr(dm2, "test_dataloader") or dm2.test_dataloader() is None:
        dm2.test_dataloader = dm2.val_dataloader
    loss = tr2.test(mod2, dm2, ckpt_path=ck)
    assert loss is not None and loss == loss and "test_loss" in tr2.callback_metrics
    for (n, a), (_, b) in zip(mod.local_named_parameters(), mod2.local_named_parameters()):
        assert torch.equal(a.detach(), b.detach()), n                   # weights came from the checkpoint
    batch = next(iter(dm2.val_dataloader()))
    outs = tr2.predict(dataloader=[batch])
    assert len(outs) == 1 and outs[0]["predictions"].shape == batch["input_ids"].shape
    assert outs[0]["logprobs"].shape == batch["input_ids"].shape and (outs[0]["logprobs"] <= 0).all()
    # greedy predictions agree with the argmax of the full logits
    with torch.no_grad():
        _, logits = mod2.model(input_ids=batch["input_ids"], labels=None)
    assert torch.equal(outs[0]["predictions"], logits.float().argmax(-1))


def test_user_callbacks_and_progress_bar(tmp_path, capsys):
    from neuronx_distributed_training_b200.trainer import Trainer
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    from neuronx_distributed_training_b200.data import SyntheticDataModule
    from neuronx_distributed_training_b200.utils.exp_manager import exp_manager
    events = []

    class Recorder:
        def on_fit_start(self, trainer, module):
            events.append("fit_start")

        def on_train_batch_end(self, trainer, module, loss, batch, batch_idx):
            events.append(("batch", trainer.global_step, float(loss) == float(loss)))

        def on_validation_end(self, trainer, module):
            events.append("val")

        def on_save_checkpoint(self, trainer, module):
            events.append("save")

        def on_train_end(self, trainer, module):
            events.append("end")

    cfg = process_config(_cfg(str(tmp_path / "run"), max_steps=4), 1)
    tr = Trainer(callbacks=[Recorder()], enable_progress_bar=True, **cfg.trainer.to_dict())
    em = exp_manager(tr, cfg.exp_manager)
    tr.fit(HFLlamaModule(cfg, tr), SyntheticDataModule(cfg, tr), ckpt_path=em.resume_from_checkpoint)
    assert events[0] == "fit_start" and events[-1] == "end"
    assert [e[1] for e in events if isinstance(e, tuple)] == [1, 2, 3, 4]
    assert events.count("val") == 2 and events.count("save") == 2          # val_check_interval = every_n_train_steps = 2
    assert "4/4" in capsys.readouterr().out


def test_save_logits_keeps_the_vocab_parallel_logits(tmp_path, monkeypatch):
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    monkeypatch.setenv("NXDT_DUMP_LOGITS_DIR", str(tmp_path / "dump"))
    cfg = process_config(_cfg(str(tmp_path / "x"), save_logits=True), 1)
    mod = HFLlamaModule(cfg, None); mod.setup(); mod.configure_optimizers(10)
    ids = torch.randint(0, 128, (2, 32))
    loss = mod.training_step({"input_ids": ids, "labels": ids.clone()})
    assert mod.last_logits.shape == (1, 32, 128) and float(loss) == float(loss)      # latest micro-batch, [B, S, V]
    assert torch.equal(torch.load(tmp_path / "dump" / "logits_rank0.pt"), mod.last_logits.cpu())


def test_fused_kernel_attachment_policy(tmp_path, monkeypatch):
    """Decision logic of BaseModelModule._maybe_enable_fused_tp / _maybe_enable_fused_zero1 with the GPU-only pieces stubbed:
    auto / true / false, environment overrides, and the modes that keep tensor parallelism on NCCL (library baseline,
    micro-batch overlap, CUDA graphs, fp8)."""
    from neuronx_distributed_training_b200 import ops
    from neuronx_distributed_training_b200.parallel import fused as fused_mod
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    cfg = _cfg(str(tmp_path / "run"))
    process_config(cfg, 1)
    mod = HFLlamaModule(cfg, None)
    mod.setup()
    assert mod.fused_tp_layers == 0                     # CPU, TP 1: nothing to attach
    calls = []
    monkeypatch.setattr(fused_mod, "enable_fused_tp", lambda model, seq_len=None, micro_batch=1: calls.append((seq_len, micro_batch)) or 7)
    # pretend to be a CUDA rank of a TP=2 sequence-parallel job
    mod.device = torch.device("cuda", 0)
    mod.dims.tp, mod.sequence_parallel = 2, True
    assert mod._maybe_enable_fused_tp() == 7 and calls[-1] == (32, 1)
    monkeypatch.setenv("NXDT_FUSED_TP", "0")
    assert mod._maybe_enable_fused_tp() == 0
    monkeypatch.delenv("NXDT_FUSED_TP")
    n = len(calls)
    for attr, val in (("use_mb_overlap", True), ("use_cuda_graph", True)):
        monkeypatch.setattr(mod, attr, val)
        assert mod._maybe_enable_fused_tp() == 0 and len(calls) == n
        monkeypatch.setattr(mod, attr, False)
    ops.set_backend("torch")
    try:
        assert mod._maybe_enable_fused_tp() == 0 and len(calls) == n
    finally:
        ops.set_backend("auto")
    ops.fp8.set_enabled(True)
    try:
        assert mod._maybe_enable_fused_tp() == 0 and len(calls) == n
    finally:
        ops.fp8.set_enabled(False)

    def boom(*a, **k):
        raise RuntimeError("no peer access")
    monkeypatch.setattr(fused_mod, "enable_fused_tp", boom)
    assert mod._maybe_enable_fused_tp() == 0            # auto: quiet fallback to NCCL
    cfg.model.fusions["fused_tp"] = True
    with pytest.raises(RuntimeError):
        mod._maybe_enable_fused_tp()                    # explicit request: the error is fatal
    cfg.model.fusions["fused_tp"] = False
    assert mod._maybe_enable_fused_tp() == 0

    # ZeRO-1: nothing on a single data-parallel rank; stubbed attach when the optimizer is sharded over 2..8 GPUs
    mod.device = torch.device("cpu")
    mod.configure_optimizers(10)
    assert mod.fused_zero1 is False
    from neuronx_distributed_training_b200.optim import fused_zero1 as fz
    attached = []
    monkeypatch.setattr(fz, "enable_fused_zero1", lambda opt: attached.append(opt) or True)
    mod.device = torch.device("cuda", 0)
    opt = mod.optimizer
    monkeypatch.setattr(opt, "zero1", True); monkeypatch.setattr(opt, "dp_size", 4)
    assert opt.master is None                           # fp32 run: no master copies -> not eligible
    assert mod._maybe_enable_fused_zero1(cfg.model.optim) is False and not attached
    monkeypatch.setattr(opt, "master", torch.zeros(1))
    assert mod._maybe_enable_fused_zero1(cfg.model.optim) is True and attached == [opt]
    monkeypatch.setenv("NXDT_FUSED_ZERO1", "0")
    assert mod._maybe_enable_fused_zero1(cfg.model.optim) is False


def test_kernels_switch_selects_the_library_path(tmp_path):
    from neuronx_distributed_training_b200 import ops
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    cfg = _cfg(str(tmp_path / "run"))
    cfg["kernels"] = "baseline"
    process_config(cfg, 1)
    assert ops.get_backend() != "torch"
    try:
        HFLlamaModule(cfg, None).setup()
        assert ops.get_backend() == "torch"
    finally:
        ops.set_backend("auto")
    cfg["kernels"] = "fast"
    with pytest.raises(ValueError):
        HFLlamaModule(cfg, None).setup()
