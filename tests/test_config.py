import glob
import os

import pytest
import torch

from neuronx_distributed_training_b200.config import (Config, load_config, from_dict, apply_overrides, process_config,
                                                      get_attribute_from_cfg, PrecisionPolicy, ParallelDims,
                                                      batch_arithmetic, padded_vocab_size)

CONF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "conf")


def test_interpolation_and_resolvers():
    cfg = from_dict({"trainer": {"max_steps": 100}, "a": {"b": 4, "c": "${.b}", "d": "${multiply:${a.b}, ${trainer.max_steps}}"},
                     "s": "steps=${trainer.max_steps}", "m": {"x": "${trainer.max_steps}"}, "lst": [1, "${a.b}"]})
    assert cfg.a.c == 4
    assert cfg.a.d == 400
    assert cfg.s == "steps=100"
    assert cfg.m.x == 100
    assert list(cfg.lst) == [1, 4]
    cfg.trainer.max_steps = 7          # lazy: overrides are visible through interpolations
    assert cfg.m.x == 7 and cfg.a.d == 28
    assert cfg.to_dict()["a"]["d"] == 28


def test_overrides_and_types():
    cfg = from_dict({"trainer": {"devices": 32}, "exp_manager": {"explicit_log_dir": None}, "model": {"optim": {"lr": 1e-4}}})
    apply_overrides(cfg, ["trainer.devices=8", "exp_manager.explicit_log_dir=null", "model.optim.lr=3e-4", "+new.key=[1,2]",
                          "flag=true"])
    assert cfg.trainer.devices == 8 and cfg.exp_manager.explicit_log_dir is None
    assert abs(cfg.model.optim.lr - 3e-4) < 1e-12 and list(cfg.new.key) == [1, 2] and cfg.flag is True
    apply_overrides(cfg, ["~flag"])
    assert "flag" not in cfg


def test_scientific_floats_parse_as_float(tmp_path):
    p = tmp_path / "c.yaml"
    p.write_text("lr: 1.5e-4\nwd: 1e-2\nname: x\n")
    cfg = load_config(str(p))
    assert isinstance(cfg.lr, float) and isinstance(cfg.wd, float)


def test_get_attribute_from_cfg_leaf_search():
    cfg = from_dict({"model": {"optim": {"lr": 0.1}, "fusions": {"flash_attention": True}}, "model_alignment_strategy": {"dpo": {"kl_beta": 0.01}}})
    assert get_attribute_from_cfg(cfg, "model.optim.lr") == 0.1
    assert get_attribute_from_cfg(cfg, "flash_attention") is True
    assert get_attribute_from_cfg(cfg, "dpo").kl_beta == 0.01
    assert get_attribute_from_cfg(cfg, "nope", 5) == 5


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(CONF, "*.yaml"))))
def test_shipped_configs_load(path):
    cfg = load_config(path)
    process_config(cfg)
    assert cfg.trainer.devices >= 1
    PrecisionPolicy.from_config(cfg)


@pytest.mark.parametrize("path", sorted(glob.glob("/root/reference/examples/conf/*.yaml")))
def test_reference_yamls_load_unmodified(path):
    cfg = load_config(path, overrides=["trainer.devices=8", "trainer.num_nodes=1"])
    process_config(cfg)
    mp = cfg.exp_manager.checkpoint_callback_params.get("model_parallel_size", None)
    ds = cfg.distributed_strategy
    if mp is not None:
        assert mp == ds.tensor_model_parallel_size * (ds.get("pipeline_model_parallel_size", 1) or 1)


def test_precision_policies():
    def pol(t, **kw):
        return PrecisionPolicy.from_config(from_dict({"precision": {"type": t, **kw}}))
    p = pol("mixed_precision")
    assert p.param_dtype == torch.bfloat16 and p.master_weights and p.fp32_grad_acc and not p.stochastic_rounding
    assert pol("mixed_precisionSR").stochastic_rounding and pol("mixed_precision_SR").stochastic_rounding
    p = pol("bf16SR")
    assert not p.master_weights and p.stochastic_rounding
    p = pol("fp32")
    assert p.param_dtype == torch.float32 and p.compute_dtype == torch.float32
    p = pol("autocast")
    assert p.param_dtype == torch.float32 and p.autocast
    p = pol("manual", master_weights=True, fp32_grad_acc=False)
    assert p.master_weights and not p.fp32_grad_acc
    assert pol("mixed_precision", parallel_layers_reduce_dtype="fp32").reduce_dtype == torch.float32
    with pytest.raises(ValueError):
        pol("int4")


def test_dims_and_batch_arithmetic():
    cfg = from_dict({"trainer": {"devices": 8, "num_nodes": 1},
                     "distributed_strategy": {"tensor_model_parallel_size": 2, "pipeline_model_parallel_size": 2},
                     "data": {"micro_batch_size": 1, "global_batch_size": 16}})
    d = ParallelDims.from_config(cfg)
    assert (d.tp, d.pp, d.dp) == (2, 2, 2)
    assert batch_arithmetic(cfg, d) == (1, 16, 8, 8)
    cfg.data.global_batch_size = 3
    with pytest.raises(ValueError):
        batch_arithmetic(cfg, d)
    assert padded_vocab_size(50257, 128, 8) == 51200


def test_moe_validation():
    base = {"model": {"activation": "swiglu", "moe": {"dropless": True, "glu_mlp": True, "capacity_factor": 2.0}}}
    cfg = process_config(from_dict(base))
    assert cfg.model.moe.capacity_factor == 0.0
    bad = from_dict({"model": {"activation": "gelu", "moe": {"dropless": True, "glu_mlp": True}}})
    with pytest.raises(ValueError):
        process_config(bad)
    bad = from_dict({"model": {"moe": {"dropless": False, "capacity_factor": 0.0}}})
    with pytest.raises(ValueError):
        process_config(bad)


def test_compile_and_train_iters_env(monkeypatch):
    cfg = from_dict({"trainer": {"max_steps": 1000}, "exp_manager": {"create_tensorboard_logger": True, "create_checkpoint_callback": True}})
    monkeypatch.setenv("COMPILE", "1")
    process_config(cfg)
    assert cfg.trainer.max_steps == 10 and not cfg.exp_manager.create_checkpoint_callback
    monkeypatch.setenv("COMPILE", "0")
    monkeypatch.setenv("TRAIN_ITERS", "3")
    cfg = process_config(from_dict({"trainer": {"max_steps": 1000}, "exp_manager": {}}))
    assert cfg.trainer.max_steps == 3
