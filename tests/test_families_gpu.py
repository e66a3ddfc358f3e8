"""Every model family / recipe takes real optimizer steps ON THE GPU through its task module, in bf16 on the native kernels
(driver-visible proof that the families are not CPU-only): Megatron-GPT variants, HF-Mixtral with the grouped expert GEMM,
LoRA-SFT, DPO, and the fp8 linear path.  The CPU twins (fp32, exact checks) live in test_megatron_gpt_cpu.py and
test_task_modules_cpu.py; here the assertions are: native kernels launched, finite losses, the loss moves the right way."""
import json

import pytest
import torch

from neuronx_distributed_training_b200 import ops
from neuronx_distributed_training_b200.parallel import state as ps

pytestmark = pytest.mark.gpu

from test_task_modules_cpu import CharTok, _cfg as _hf_cfg, _fit, _pref_file          # noqa: E402
from test_megatron_gpt_cpu import _cfg as _mega_cfg, _module as _mega_module             # noqa: E402


@pytest.fixture(autouse=True)
def _mp():
    ps.initialize_model_parallel(1, 1, 1)
    yield
    ps.destroy_model_parallel()


GPU_MODEL = {"encoder_seq_length": 256, "max_position_embeddings": 256, "hidden_size": 256, "intermediate_size": 512,
             "num_attention_heads": 2, "num_kv_heads": 1, "vocab_size": 512, "fusions": {"flash_attention": True}}
BF16 = {"precision": {"type": "mixed_precision"}}


@pytest.mark.parametrize("name,kw", [
    ("gpt2_style", {}),
    ("llama_style", {"normalization": "rmsnorm", "position_embedding_type": "rope", "activation": "swiglu", "has_bias": False,
                     "share_embeddings_and_output_weights": False}),
    ("moe_top2", {"moe": {"num_experts": 4, "top_k": 2, "frequency": 1, "routing_algorithm": "top_k", "glu_mlp": True,
                          "router_aux_loss_coef": 0.01, "dropless": True}, "activation": "swiglu", "has_bias": False}),
])
def test_megatron_gpt_trains_on_gpu(name, kw):
    cfg = _mega_cfg({"encoder_seq_length": 128, "max_position_embeddings": 128, "hidden_size": 256, "ffn_hidden_size": 512,
                     "num_attention_heads": 4, "vocab_size": 512, **kw}, data_kw={"seq_length": 128}, top_kw=BF16)
    mod = _mega_module(cfg)
    assert mod.device.type == "cuda"
    g = torch.Generator().manual_seed(0)
    toks = torch.randint(0, 512, (2, 129), generator=g)
    batch = {"tokens": toks[:, :-1].contiguous(), "labels": toks[:, 1:].contiguous(), "loss_mask": torch.ones(2, 128),
             "position_ids": torch.arange(128).unsqueeze(0).expand(2, 128).contiguous()}
    ops.reset_native_launch_count()
    losses = [float(mod.training_step(dict(batch))) for _ in range(8)]
    assert ops.native_launch_count() > 0, "no native kernel ran"
    assert all(l == l for l in losses) and losses[-1] < losses[0] - 0.2, (name, losses)


def test_mixtral_dropless_trains_on_gpu_with_the_grouped_expert_gemm(tmp_path):
    from neuronx_distributed_training_b200.trainer.hf_mixtral_module import HFMixtralModule
    from neuronx_distributed_training_b200.data import SyntheticDataModule
    cfg = _hf_cfg(str(tmp_path / "run"), name="hf_mixtral", max_steps=6,
                  model_kw={**GPU_MODEL, "model_config": "mixtral_8x7b", "sliding_window": 128,
                            "moe": {"num_experts": 4, "top_k": 2, "dropless": True, "router_aux_loss_coef": 0.02}},
                  top_kw=BF16)
    ops.reset_native_launch_count()
    tr, mod, losses = _fit(cfg, HFMixtralModule, SyntheticDataModule)
    assert mod.device.type == "cuda" and tr.global_step == 6 and all(l == l for l in losses)
    assert ops.native_launch_count() > 0
    from neuronx_distributed_training_b200.models import moe
    assert getattr(moe, "_GROUPED_CALLS", {"n": 1})["n"] > 0, "the grouped expert GEMM did not run"


def test_dpo_trains_on_gpu(tmp_path):
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    from neuronx_distributed_training_b200.data.alignment import ModelAlignmentDataModule
    cfg = _hf_cfg(str(tmp_path / "run"), max_steps=5, model_kw=GPU_MODEL,
                  data_kw={"synthetic": False, "train_dir": _pref_file(tmp_path / "p.jsonl"), "seq_length": 256},
                  top_kw={**BF16, "model_alignment_strategy": {"dpo": {"kl_beta": 0.1, "max_prompt_length": 64,
                                                                       "precompute_ref_log_probs": True,
                                                                       "truncation_mode": "keep_start"}}})
    tr, mod, losses = _fit(cfg, HFLlamaModule, lambda c, t: ModelAlignmentDataModule(c, t, tokenizer=CharTok()))
    assert mod.device.type == "cuda"
    assert losses[0] == pytest.approx(0.6931, abs=2e-2)           # the first step sees the reference policy itself
    assert all(l == l for l in losses) and losses[-1] < losses[0]


def test_lora_sft_trains_on_gpu(tmp_path):
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    from neuronx_distributed_training_b200.data.alignment import ModelAlignmentDataModule
    p = tmp_path / "sft.jsonl"
    p.write_text("\n".join(json.dumps({"input": "say " + "q" * (i % 4), "output": "ok " + "a" * (i % 6)}) for i in range(16)))
    peft = {"lora_rank": 8, "lora_alpha": 16, "lora_dropout": 0.0, "lora_bias": "none",
            "target_modules": ["qkv_proj", "o_proj", "gate_up_proj", "down_proj"]}
    cfg = _hf_cfg(str(tmp_path / "run"), max_steps=6, model_kw=GPU_MODEL,
                  data_kw={"synthetic": False, "train_dir": str(p), "seq_length": 256},
                  top_kw={"model_alignment_strategy": {"sft": {"packing": False}, "peft": peft},
                          "distributed_strategy": {"tensor_model_parallel_size": 1, "zero1": False},
                          "precision": {"type": "autocast"}})
    tr, mod, losses = _fit(cfg, HFLlamaModule, lambda c, t: ModelAlignmentDataModule(c, t, tokenizer=CharTok()))
    assert mod.device.type == "cuda" and all(l == l for l in losses) and losses[-1] < losses[0]
    trainable = [n for n, q in mod.model.named_parameters() if q.requires_grad]
    assert trainable and all("lora" in n.lower() for n in trainable)


def test_fp8_linear_path_trains_on_gpu(tmp_path):
    """model.fusions.fp8_gemm: every linear GEMM on the e4m3 tcgen05 kernel (kind::f8f6f4, row / column block scales)."""
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    from neuronx_distributed_training_b200.data import SyntheticDataModule
    from neuronx_distributed_training_b200.ops import fp8
    cfg = _hf_cfg(str(tmp_path / "run"), max_steps=6, model_kw={**GPU_MODEL, "fusions": {"flash_attention": True, "fp8_gemm": True}},
                  data_kw={"num_samples": 2}, top_kw=BF16)
    try:
        tr, mod, losses = _fit(cfg, HFLlamaModule, SyntheticDataModule)
        assert fp8.enabled()
    finally:
        fp8.set_enabled(False)
    assert all(l == l for l in losses) and losses[-1] < losses[0]
