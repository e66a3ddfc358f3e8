"""Megatron-GPT model family on CPU: architecture variants train, config validation, HF-style (SFT) batches."""
import pytest
import torch

from neuronx_distributed_training_b200.config import from_dict, process_config
from neuronx_distributed_training_b200.parallel import state as ps


@pytest.fixture(autouse=True)
def _mp():
    ps.initialize_model_parallel(1, 1, 1)
    yield
    ps.destroy_model_parallel()


def _cfg(model_kw=None, data_kw=None, top_kw=None, ds_kw=None):
    model = {"encoder_seq_length": 32, "max_position_embeddings": 32, "num_layers": 2, "hidden_size": 64, "ffn_hidden_size": 128,
             "num_attention_heads": 4, "vocab_size": 128, "init_method_std": 0.05, "hidden_dropout": 0.0, "attention_dropout": 0.0,
             "normalization": "layernorm", "position_embedding_type": "learned_absolute", "activation": "gelu", "has_bias": True,
             "share_embeddings_and_output_weights": True, "make_vocab_size_divisible_by": 8,
             "fusions": {"flash_attention": False},
             "optim": {"name": "adamw", "lr": 2e-2, "weight_decay": 0.0, "betas": [0.9, 0.95],
                       "sched": {"name": "CosineAnnealing", "warmup_steps": 1, "constant_steps": 0, "min_lr": 1e-3, "max_steps": 50}}}
    model.update(model_kw or {})
    d = {"name": "megatron_gpt", "model_source": "megatron", "seed": 5,
         "trainer": {"devices": 1, "num_nodes": 1, "max_steps": 50, "gradient_clip_val": 1.0},
         "exp_manager": {"create_checkpoint_callback": False, "create_tensorboard_logger": False},
         "distributed_strategy": {"tensor_model_parallel_size": 1, "zero1": True, **(ds_kw or {})},
         "data": {"micro_batch_size": 2, "global_batch_size": 2, "seq_length": 32, **(data_kw or {})},
         "model": model, "precision": {"type": "fp32"}}
    d.update(top_kw or {})
    return process_config(from_dict(d), 1)


def _module(cfg):
    from neuronx_distributed_training_b200.trainer.megatron_module import MegatronGPTModel
    mod = MegatronGPTModel(cfg, None)
    mod.setup()
    mod.configure_optimizers(50)
    return mod


def _megatron_batch(seed=0, B=2, S=32, V=128):
    g = torch.Generator().manual_seed(seed)
    toks = torch.randint(0, V, (B, S + 1), generator=g)
    return {"tokens": toks[:, :-1].contiguous(), "labels": toks[:, 1:].contiguous(),
            "loss_mask": torch.ones(B, S), "position_ids": torch.arange(S).unsqueeze(0).expand(B, S).contiguous()}


VARIANTS = {
    "gpt2_style": {},
    "llama_style": {"normalization": "rmsnorm", "position_embedding_type": "rope", "activation": "swiglu", "has_bias": False,
                    "share_embeddings_and_output_weights": False},
    "gqa_rope_partial": {"num_kv_heads": 2, "position_embedding_type": "rope", "rotary_percentage": 0.5, "activation": "geglu"},
    "layernorm1p_reglu": {"normalization": "layernorm1p", "activation": "reglu"},
    "post_ln": {"transformer_block_type": "post_ln"},
    "normformer": {"transformer_block_type": "normformer"},
    "gpt_j": {"transformer_block_type": "gpt_j", "position_embedding_type": "rope"},
    "sliding_window": {"sliding_window_size": 8, "position_embedding_type": "rope", "normalization": "rmsnorm"},
    "moe_top2": {"moe": {"num_experts": 4, "top_k": 2, "frequency": 1, "routing_algorithm": "top_k", "glu_mlp": True,
                         "router_aux_loss_coef": 0.01, "dropless": True}, "activation": "swiglu", "has_bias": False},
    "moe_sinkhorn": {"moe": {"num_experts": 4, "top_k": 1, "routing_algorithm": "sinkhorn", "glu_mlp": False, "capacity_factor": 2.0},
                     "activation": "gelu"},
    "dropout": {"hidden_dropout": 0.1, "attention_dropout": 0.1, "ffn_dropout": 0.1},
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_variant_trains(name):
    mod = _module(_cfg(VARIANTS[name]))
    batch = _megatron_batch()
    losses = [float(mod.training_step(dict(batch))) for _ in range(8)]
    assert all(l == l for l in losses), losses                          # finite
    assert losses[-1] < losses[0] - 0.2, (name, losses)                   # memorises the batch
    # state-dict names follow the Megatron convention the converters rely on
    keys = list(mod.state_dict().keys())
    assert any(k.startswith("language_model.embedding.word_embeddings") for k in keys)
    assert any(".self_attention." in k for k in keys)
    assert any("final_layernorm" in k for k in keys) == (name != "post_ln")          # post-LN blocks end normalised


def test_loss_mask_removes_tokens_from_the_loss():
    mod = _module(_cfg())
    b = _megatron_batch()
    with torch.no_grad():
        full, _ = mod.model_fwd_calc_loss(b)
        b2 = dict(b)
        b2["loss_mask"] = b["loss_mask"].clone()
        b2["loss_mask"][:, 16:] = 0
        half, _ = mod.model_fwd_calc_loss(b2)
        b3 = dict(b)
        b3["labels"] = b["labels"].clone()
        b3["labels"][:, 16:] = 0                                          # masked positions must not matter
        b3["loss_mask"] = b2["loss_mask"]
        half2, _ = mod.model_fwd_calc_loss(b3)
    assert abs(float(full) - float(half)) > 1e-6
    assert float(half) == pytest.approx(float(half2), abs=1e-6)


def test_hf_style_sft_batches_are_shifted():
    """data.use_sft_style_data_module: labels come aligned with the inputs (-100 on the prompt), the module shifts them."""
    torch.manual_seed(0)
    ids = torch.randint(0, 128, (2, 32))
    labels = ids.clone()
    labels[:, :10] = -100
    mod = _module(_cfg(data_kw={"use_sft_style_data_module": True}))
    with torch.no_grad():
        loss_sft, _ = mod.model_fwd_calc_loss({"input_ids": ids, "labels": labels})
        shifted = torch.cat([labels[:, 1:], torch.full((2, 1), -100)], dim=1)
        mask = (shifted != -100).float()
        loss_ref, _ = mod.model_fwd_calc_loss({"tokens": ids, "labels": shifted.clamp_min(0), "loss_mask": mask})
    assert float(loss_sft) == pytest.approx(float(loss_ref), abs=1e-6)


def test_config_validation():
    from neuronx_distributed_training_b200.trainer.megatron_module import MegatronGPTModel
    with pytest.raises(ValueError):
        MegatronGPTModel(_cfg(ds_kw={"zero1": False}), None)
    with pytest.raises(ValueError):
        MegatronGPTModel(_cfg({"optim": {"name": "distributed_fused_adam", "lr": 1e-3}}), None)
    with pytest.raises(ValueError):
        MegatronGPTModel(_cfg({"megatron_amp_O2": True}), None)
    with pytest.raises(ValueError):
        MegatronGPTModel(_cfg({"activations_checkpoint_granularity": "full", "activations_checkpoint_num_layers": 2}), None)
    with pytest.raises(ValueError):
        _module(_cfg({"activations_checkpoint_granularity": "selective", "activations_checkpoint_recompute": ["Nope"]}))


def test_gpt2_style_model_matches_transformers_gpt2():
    """Independent oracle for the Megatron-GPT path (learned positions, LayerNorm, biases, tanh-GeLU, tied embeddings,
    per-head interleaved QKV): the weights of a `transformers` GPT2LMHeadModel give the same loss."""
    tr = pytest.importorskip("transformers")
    torch.manual_seed(0)
    H, nh, L, S, V, F_ = 64, 4, 2, 32, 128, 128
    hd = H // nh
    hcfg = tr.GPT2Config(vocab_size=V, n_positions=S, n_embd=H, n_layer=L, n_head=nh, n_inner=F_, activation_function="gelu_new",
                         resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0, layer_norm_epsilon=1e-5, attn_implementation="eager")
    hf = tr.GPT2LMHeadModel(hcfg).float().eval()
    with torch.no_grad():
        for p in hf.parameters():
            p.normal_(0, 0.05)                         # non-trivial biases / norms (the LM head is tied to wte)
    mod = _module(_cfg({"layernorm_epsilon": 1e-5}))
    sd = {k: v.detach() for k, v in hf.state_dict().items()}
    new = {"language_model.embedding.word_embeddings.weight": sd["transformer.wte.weight"],
           "language_model.embedding.position_embeddings.weight": sd["transformer.wpe.weight"],
           "language_model.encoder.final_layernorm.weight": sd["transformer.ln_f.weight"],
           "language_model.encoder.final_layernorm.bias": sd["transformer.ln_f.bias"]}
    for l in range(L):
        a, b = f"transformer.h.{l}", f"language_model.encoder.layers.{l}"
        for src, dst in (("ln_1", "input_layernorm"), ("ln_2", "post_attention_layernorm")):
            new[f"{b}.{dst}.weight"], new[f"{b}.{dst}.bias"] = sd[f"{a}.{src}.weight"], sd[f"{a}.{src}.bias"]
        w = sd[f"{a}.attn.c_attn.weight"].t()          # Conv1D stores [in, out]: rows of w are [q_all; k_all; v_all]
        bias = sd[f"{a}.attn.c_attn.bias"]
        wq, wk, wv = w.view(3, nh, hd, H)
        bq, bk, bv = bias.view(3, nh, hd)
        new[f"{b}.self_attention.query_key_value.weight"] = torch.stack([wq, wk, wv], 1).reshape(3 * H, H)     # per head: q, k, v
        new[f"{b}.self_attention.query_key_value.bias"] = torch.stack([bq, bk, bv], 1).reshape(3 * H)
        new[f"{b}.self_attention.dense.weight"] = sd[f"{a}.attn.c_proj.weight"].t()
        new[f"{b}.self_attention.dense.bias"] = sd[f"{a}.attn.c_proj.bias"]
        new[f"{b}.mlp.dense_h_to_4h.weight"] = sd[f"{a}.mlp.c_fc.weight"].t()
        new[f"{b}.mlp.dense_h_to_4h.bias"] = sd[f"{a}.mlp.c_fc.bias"]
        new[f"{b}.mlp.dense_4h_to_h.weight"] = sd[f"{a}.mlp.c_proj.weight"].t()
        new[f"{b}.mlp.dense_4h_to_h.bias"] = sd[f"{a}.mlp.c_proj.bias"]
    params = {n.split("chunks.0.", 1)[-1]: p for n, p in mod.model.named_parameters()}
    assert set(params) == set(new), sorted(set(params) ^ set(new))[:6]
    with torch.no_grad():
        for n, p in params.items():
            p.copy_(new[n].reshape(p.shape))
    ids = torch.randint(0, V, (2, S))
    with torch.no_grad():
        want = hf(input_ids=ids, labels=ids).loss
        labels = torch.cat([ids[:, 1:], ids[:, :1]], 1)
        mask = torch.ones(2, S); mask[:, -1] = 0
        got, _ = mod.model_fwd_calc_loss({"tokens": ids, "labels": labels, "loss_mask": mask,
                                          "position_ids": torch.arange(S).unsqueeze(0).expand(2, S).contiguous()})
    assert abs(float(got) - float(want)) < 2e-5, (float(got), float(want))


def test_llama_style_megatron_model_matches_transformers_llama():
    """RoPE + RMSNorm + SwiGLU + untied head on the Megatron path against `transformers` LlamaForCausalLM (MHA)."""
    tr = pytest.importorskip("transformers")
    torch.manual_seed(0)
    H, nh, L, S, V, F_ = 64, 4, 2, 32, 128, 128
    hd = H // nh
    hcfg = tr.LlamaConfig(vocab_size=V, hidden_size=H, intermediate_size=F_, num_hidden_layers=L, num_attention_heads=nh,
                          num_key_value_heads=nh, max_position_embeddings=S, rope_theta=10000.0, rms_norm_eps=1e-5,
                          tie_word_embeddings=False, attn_implementation="eager")
    hf = tr.LlamaForCausalLM(hcfg).float().eval()
    sd = {k: v.detach() for k, v in hf.state_dict().items()}
    mod = _module(_cfg({**VARIANTS["llama_style"], "layernorm_epsilon": 1e-5}))
    new = {"language_model.embedding.word_embeddings.weight": sd["model.embed_tokens.weight"],
           "language_model.output_layer.weight": sd["lm_head.weight"],
           "language_model.encoder.final_layernorm.weight": sd["model.norm.weight"]}
    for l in range(L):
        a, b = f"model.layers.{l}", f"language_model.encoder.layers.{l}"
        new[f"{b}.input_layernorm.weight"] = sd[f"{a}.input_layernorm.weight"]
        new[f"{b}.post_attention_layernorm.weight"] = sd[f"{a}.post_attention_layernorm.weight"]
        wq, wk, wv = (sd[f"{a}.self_attn.{n}_proj.weight"].view(nh, hd, H) for n in "qkv")
        new[f"{b}.self_attention.query_key_value.weight"] = torch.stack([wq, wk, wv], 1).reshape(3 * H, H)
        new[f"{b}.self_attention.dense.weight"] = sd[f"{a}.self_attn.o_proj.weight"]
        new[f"{b}.mlp.dense_h_to_4h.weight"] = torch.cat([sd[f"{a}.mlp.gate_proj.weight"], sd[f"{a}.mlp.up_proj.weight"]], 0)
        new[f"{b}.mlp.dense_4h_to_h.weight"] = sd[f"{a}.mlp.down_proj.weight"]
    params = {n.split("chunks.0.", 1)[-1]: p for n, p in mod.model.named_parameters()}
    assert set(params) == set(new), sorted(set(params) ^ set(new))[:6]
    with torch.no_grad():
        for n, p in params.items():
            p.copy_(new[n].reshape(p.shape))
    ids = torch.randint(0, V, (2, S))
    with torch.no_grad():
        want = hf(input_ids=ids, labels=ids).loss
        labels = torch.cat([ids[:, 1:], ids[:, :1]], 1)
        mask = torch.ones(2, S); mask[:, -1] = 0
        got, _ = mod.model_fwd_calc_loss({"tokens": ids, "labels": labels, "loss_mask": mask,
                                          "position_ids": torch.arange(S).unsqueeze(0).expand(2, S).contiguous()})
    assert abs(float(got) - float(want)) < 2e-5, (float(got), float(want))
