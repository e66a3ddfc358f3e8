"""End-to-end runs of the task modules through Trainer.fit on CPU: Mixtral (dropless / capacity), DPO, ORPO, LoRA-SFT."""
import json
import os

import pytest
import torch

from neuronx_distributed_training_b200.config import from_dict, process_config
from neuronx_distributed_training_b200.parallel import state as ps


class CharTok:
    bos_token_id, eos_token_id, pad_token_id = 1, 2, 0

    def encode(self, text, add_special_tokens=False):
        return [3 + (ord(c) % 60) for c in text]


@pytest.fixture(autouse=True)
def _mp():
    ps.initialize_model_parallel(1, 1, 1)
    yield
    ps.destroy_model_parallel()
    import logging
    lg = logging.getLogger("nxdt")
    for h in list(lg.handlers):
        lg.removeHandler(h)


def _cfg(log_dir, name="hf_llama", max_steps=6, model_kw=None, data_kw=None, top_kw=None):
    model = {"model_config": "llama_125m", "encoder_seq_length": 32, "max_position_embeddings": 32, "num_layers": 2, "hidden_size": 64,
             "intermediate_size": 128, "num_attention_heads": 4, "num_kv_heads": 2, "vocab_size": 128,
             "fusions": {"flash_attention": False},
             "optim": {"name": "adamw_fp32OptState", "lr": 2e-2, "weight_decay": 0.0, "betas": [0.9, 0.95],
                       "sched": {"name": "LinearAnnealingWithWarmUp", "warmup_steps": 1, "max_steps": 50}}}
    model.update(model_kw or {})
    d = {"name": name, "model_source": "hf", "seed": 3,
         "trainer": {"devices": 1, "num_nodes": 1, "max_steps": max_steps, "log_every_n_steps": 1, "gradient_clip_val": 1.0},
         "exp_manager": {"explicit_log_dir": log_dir, "create_tensorboard_logger": False, "create_checkpoint_callback": False},
         "distributed_strategy": {"tensor_model_parallel_size": 1, "zero1": True},
         "data": {"micro_batch_size": 1, "global_batch_size": 2, "synthetic": True, **(data_kw or {})},
         "model": model, "precision": {"type": "fp32"}}
    d.update(top_kw or {})
    return from_dict(d)


def _fit(cfg, module_cls, datamodule):
    from neuronx_distributed_training_b200.trainer import Trainer
    from neuronx_distributed_training_b200.utils.exp_manager import exp_manager
    process_config(cfg, 1)
    tr = Trainer(**cfg.trainer.to_dict())
    em = exp_manager(tr, cfg.exp_manager)
    mod = module_cls(cfg, tr)
    losses = []
    orig = mod.training_step

    def wrapped(b, i=0):
        l = orig(b, i)
        losses.append(float(l))
        return l
    mod.training_step = wrapped
    dm = datamodule(cfg, tr) if callable(datamodule) else datamodule
    tr.fit(mod, dm, ckpt_path=em.resume_from_checkpoint)
    return tr, mod, losses


@pytest.mark.parametrize("moe", [{"num_experts": 4, "top_k": 2, "dropless": True, "router_aux_loss_coef": 0.02},
                                 {"num_experts": 4, "top_k": 2, "capacity_factor": 2.0, "router_aux_loss_coef": 0.02},
                                 {"num_experts": 4, "top_k": 1, "capacity_factor": 4.0, "frequency": 2}])
def test_mixtral_module_trains(tmp_path, moe):
    from neuronx_distributed_training_b200.trainer.hf_mixtral_module import HFMixtralModule
    from neuronx_distributed_training_b200.data import SyntheticDataModule
    cfg = _cfg(str(tmp_path / "run"), name="hf_mixtral", max_steps=8,
               model_kw={"model_config": "mixtral_8x7b", "moe": moe, "sliding_window": 16})
    tr, mod, losses = _fit(cfg, HFMixtralModule, SyntheticDataModule)
    assert tr.global_step == 8 and all(l == l for l in losses)
    keys = list(mod.state_dict().keys())
    assert any("expert_mlps.mlp_op.gate_up_proj" in k for k in keys) and any("router" in k for k in keys)
    # the synthetic stream is random tokens: the loss must at least move towards log(V) from its initial value
    assert abs(losses[-1] - losses[0]) > 1e-4


def _pref_file(path, n=12):
    recs = [{"prompt": "question " + "x" * (i % 3), "chosen": "good answer " + "y" * (i % 4), "rejected": "bad" + "z" * (i % 5)}
            for i in range(n)]
    path.write_text("\n".join(json.dumps(r) for r in recs))
    return str(path)


def test_dpo_end_to_end(tmp_path):
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    from neuronx_distributed_training_b200.data.alignment import ModelAlignmentDataModule
    cfg = _cfg(str(tmp_path / "run"), max_steps=5,
               data_kw={"synthetic": False, "train_dir": _pref_file(tmp_path / "p.jsonl"), "seq_length": 32},
               top_kw={"model_alignment_strategy": {"dpo": {"kl_beta": 0.1, "max_prompt_length": 16,
                                                            "precompute_ref_log_probs": True, "truncation_mode": "keep_start"}}})
    tr, mod, losses = _fit(cfg, HFLlamaModule, lambda c, t: ModelAlignmentDataModule(c, t, tokenizer=CharTok()))
    ds = tr.datamodule.train_dataset()
    assert "reference_chosen_logps" in ds.extra and len(ds.extra["reference_chosen_logps"]) == len(ds)
    # the first step sees the reference policy itself: loss = -log sigmoid(0) = log 2, rewards = 0
    assert losses[0] == pytest.approx(0.6931, abs=2e-3)
    assert losses[-1] < losses[0]
    m = tr.logged_metrics
    assert "chosen_rewards" in m and "rejected_rewards" in m and m["chosen_rewards"] > m["rejected_rewards"]


def test_orpo_end_to_end(tmp_path):
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    from neuronx_distributed_training_b200.data.alignment import ModelAlignmentDataModule
    cfg = _cfg(str(tmp_path / "run"), max_steps=6,
               data_kw={"synthetic": False, "train_dir": _pref_file(tmp_path / "p.jsonl"), "seq_length": 32},
               top_kw={"model_alignment_strategy": {"orpo": {"beta": 0.2, "max_prompt_length": 16, "truncation_mode": "keep_start"}}})
    tr, mod, losses = _fit(cfg, HFLlamaModule, lambda c, t: ModelAlignmentDataModule(c, t, tokenizer=CharTok()))
    assert tr.global_step == 6 and all(l == l for l in losses) and losses[-1] < losses[0]


def test_lora_sft_trains_only_the_adapters(tmp_path):
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    from neuronx_distributed_training_b200.data.alignment import ModelAlignmentDataModule
    p = tmp_path / "sft.jsonl"
    p.write_text("\n".join(json.dumps({"input": "say " + "q" * (i % 4), "output": "ok " + "a" * (i % 6)}) for i in range(16)))
    peft = {"lora_rank": 4, "lora_alpha": 8, "lora_dropout": 0.0, "lora_bias": "none",
            "target_modules": ["qkv_proj", "o_proj", "gate_up_proj", "down_proj"]}
    cfg = _cfg(str(tmp_path / "run"), max_steps=6,
               data_kw={"synthetic": False, "train_dir": str(p), "seq_length": 32},
               top_kw={"model_alignment_strategy": {"sft": {"packing": False}, "peft": peft},
                       "distributed_strategy": {"tensor_model_parallel_size": 1, "zero1": False},
                       "precision": {"type": "autocast"}})
    from neuronx_distributed_training_b200.trainer import Trainer
    from neuronx_distributed_training_b200.utils.exp_manager import exp_manager
    process_config(cfg, 1)
    tr = Trainer(**cfg.trainer.to_dict())
    exp_manager(tr, cfg.exp_manager)
    mod = HFLlamaModule(cfg, tr)
    mod.setup()
    base_before = {n: p_.detach().clone() for n, p_ in mod.model.named_parameters() if "adapter" not in n and "lora" not in n.lower()}
    trainable = [n for n, p_ in mod.model.named_parameters() if p_.requires_grad]
    assert trainable and all(("adapter" in n or "lora" in n.lower()) for n in trainable), trainable[:5]
    tr.fit(mod, ModelAlignmentDataModule(cfg, tr, tokenizer=CharTok()))
    assert tr.global_step == 6
    for n, p_ in mod.model.named_parameters():
        if n in base_before:
            assert torch.equal(p_.detach(), base_before[n]), n                     # frozen base
    sd = mod.state_dict()
    assert sd and all(("adapter" in k or "lora" in k.lower()) for k in sd), list(sd)[:4]   # adapter-only checkpoint by default
    # save_lora_config_adapter (default on): the adapter description is stored next to the adapter weights
    ck = str(tmp_path / "manual" / "lora--step=6-consumed_samples=0.ckpt")
    tr.save_checkpoint(ck)
    with open(os.path.join(ck, "adapter_config.json")) as f:
        ac = json.load(f)
    assert ac["peft_type"] == "LORA" and ac["r"] == 4 and ac["lora_alpha"] == 8 and ac["target_modules"] == peft["target_modules"]
    assert os.path.exists(os.path.join(ck, "done")) and os.path.exists(os.path.join(ck, "user_content.pt"))
    uc = torch.load(os.path.join(ck, "user_content.pt"), weights_only=False)
    assert "_extra_files" not in uc
