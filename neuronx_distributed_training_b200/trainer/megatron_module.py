"""Megatron-GPT task module: YAML -> GPTConfig mapping + config validation.

Parity with the reference's ``MegatronBaseModel`` / ``MegatronGPTModel`` (reference
megatron/megatron_base_model.py:28-129, megatron_gpt_model.py:61-259): SP forced off at TP=1, ZeRO-1 required,
rejected options (distributed fused adam, amp O2, grad-accumulation fusion, bias fusions with MoE dropout ...),
~60 YAML keys mapped onto the model, batches unpacked as (tokens, labels, loss_mask, position_ids).
"""
from __future__ import annotations

import logging
from typing import Dict

import torch

from ..config import Config, padded_vocab_size
from ..models import megatron_gpt as gpt_mod
from ..models.megatron_gpt import GPTConfig, GPTModel, ParallelTransformerLayer
from ..parallel.model import initialize_parallel_model
from .module import BaseModelModule, get_param_groups_by_weight_decay

log = logging.getLogger("nxdt")


class MegatronBaseModel(BaseModelModule):
    def __init__(self, cfg: Config, trainer=None):
        super().__init__(cfg, trainer)
        self._validate()

    def _validate(self):
        m, ds = self.config, self.cfg.get("distributed_strategy", Config({}))
        if not bool(ds.get("zero1", True)):
            raise ValueError("Megatron models require distributed_strategy.zero1=True (reference megatron_base_model.py:85-86)")
        if str(m.optim.get("name", "")) == "distributed_fused_adam":
            raise ValueError("distributed_fused_adam is not supported; use adamw / adamw_fp32OptState with zero1")
        if bool(m.get("megatron_amp_O2", False)):
            raise ValueError("megatron_amp_O2 is not supported; choose a precision.type instead")
        if bool(m.get("gradient_accumulation_fusion", False)):
            log.info("gradient_accumulation_fusion is always on here (wgrad GEMMs accumulate into fp32 main_grad)")
        if m.get("activations_checkpoint_granularity", None) == "full" and int(m.get("activations_checkpoint_num_layers", 1) or 1) > 1:
            raise ValueError("full activation checkpointing supports activations_checkpoint_num_layers=1 only")
        vpp = int(ds.get("virtual_pipeline_model_parallel_size", 1) or 1)
        if vpp > 1 and int(m.num_layers) % (self.dims.pp * vpp) != 0:
            raise ValueError("num_layers must be divisible by pipeline_model_parallel_size * virtual_pipeline size")
        moe = m.get("moe", None)
        if moe is not None and float(moe.get("dropout", 0.0) or 0.0) > 0.0:
            raise ValueError("MoE dropout is not supported")

    def setup_optimizer_param_groups(self):
        wd = float(self.config.optim.get("weight_decay", 0.01))
        return get_param_groups_by_weight_decay(self.local_named_parameters(), wd, no_decay_norm=True)


class MegatronGPTModel(MegatronBaseModel):
    def _gpt_config(self) -> GPTConfig:
        m = self.config
        tp = self.dims.tp
        vocab = m.get("vocab_size", None)
        if vocab is None:
            tok = self.cfg.data.get("tokenizer", Config({})) if "data" in self.cfg else Config({})
            vocab = int(tok.get("vocab_size", 50257))
        vocab = padded_vocab_size(int(vocab), int(m.get("make_vocab_size_divisible_by", 128)), tp)
        moe = m.get("moe", Config({})) or Config({})
        ds = self.cfg.get("distributed_strategy", Config({}))
        fus = m.get("fusions", Config({})) or Config({})
        return GPTConfig(
            vocab_size=vocab, hidden_size=int(m.hidden_size), ffn_hidden_size=int(m.get("ffn_hidden_size", 4 * int(m.hidden_size))),
            num_layers=int(m.num_layers), num_attention_heads=int(m.num_attention_heads),
            num_kv_heads=m.get("num_kv_heads", None) or m.get("num_query_groups", None),
            kv_channels=m.get("kv_channels", None), max_position_embeddings=int(m.get("max_position_embeddings", 2048)),
            encoder_seq_length=int(m.get("encoder_seq_length", 2048)), init_method_std=float(m.get("init_method_std", 0.02)),
            use_scaled_init_method=bool(m.get("use_scaled_init_method", True)),
            hidden_dropout=float(m.get("hidden_dropout", 0.1)), attention_dropout=float(m.get("attention_dropout", 0.1)),
            ffn_dropout=float(m.get("ffn_dropout", 0.0)),
            apply_query_key_layer_scaling=bool(m.get("apply_query_key_layer_scaling", False)),
            normalization=str(m.get("normalization", "layernorm")).lower(), layernorm_epsilon=float(m.get("layernorm_epsilon", 1e-5)),
            share_embeddings_and_output_weights=bool(m.get("share_embeddings_and_output_weights", True)),
            position_embedding_type=str(m.get("position_embedding_type", "learned_absolute")),
            rotary_percentage=float(m.get("rotary_percentage", 1.0)), rotary_base=float(m.get("rotary_base", 10000.0)),
            position_interpolation_factor=float(m.get("position_interpolation_factor", 1.0) or 1.0),
            position_abf_factor=float(m.get("position_abf_factor", 1.0) or 1.0),
            activation=str(m.get("activation", "gelu")), transformer_block_type=str(m.get("transformer_block_type", "pre_ln")),
            has_bias=bool(m.get("has_bias", m.get("bias", True))), sliding_window_size=m.get("sliding_window_size", None),
            sequence_parallel=self.sequence_parallel, use_flash_attention=bool(fus.get("flash_attention", True)),
            params_dtype=self.precision.param_dtype, use_cpu_initialization=not torch.cuda.is_available(),
            num_moe_experts=int(moe.get("num_experts", 1) or 1), moe_frequency=int(moe.get("frequency", 1) or 1),
            moe_top_k=int(moe.get("top_k", 2)), moe_routing_algorithm=str(moe.get("routing_algorithm", "top_k")),
            moe_router_activation=str(moe.get("router_activation", "softmax")),
            moe_capacity_factor=float(moe.get("capacity_factor", 0.0) or 0.0), moe_dropless=bool(moe.get("dropless", False)),
            moe_sinkhorn_iterations=int(moe.get("sinkhorn_iterations", 30)), moe_sinkhorn_tol=moe.get("sinkhorn_tol", None),
            moe_aux_loss_coef=float(moe.get("router_aux_loss_coef", 0.0) or 0.0) if bool(moe.get("output_router_logits", True)) else 0.0,
            moe_normalize_top_k_affinities=bool(moe.get("normalize_top_k_affinities", True)), moe_glu_mlp=bool(moe.get("glu_mlp", True)),
            expert_model_parallel_size=int(ds.get("expert_model_parallel_size", 1) or 1),
            token_shuffle_group_size=int(ds.get("token_shuffle_group_size", 1) or 1),
        )

    def _activation_checkpoint_spec(self):
        m = self.config
        gran = m.get("activations_checkpoint_granularity", None)
        if gran == "full":
            return "full"
        if gran == "selective":
            names = list(m.get("activations_checkpoint_recompute", None) or ["CoreAttention"])
            for n in names:
                if not (hasattr(gpt_mod, n) or n == "NeuronSwitchMLP"):
                    raise ValueError(f"unknown recompute class {n}")
            return names
        return None

    def build_model(self):
        g = self._gpt_config()
        self.gpt_config = g
        if self.dims.cp > 1:
            # the reference offers context parallelism for HF-Llama only (ring attention lives in modeling_llama.py:483-484);
            # the Megatron attention has no ring path and its datasets hand out pre-shifted labels, so a sequence split
            # here would silently train chunk-local attention on doubly shifted labels
            raise ValueError("context_parallel_size > 1 is not supported for Megatron-GPT models (HF-Llama only); "
                             "set distributed_strategy.context_parallel_size=1")
        if self.dims.pp > 1 and g.share_embeddings_and_output_weights:
            raise ValueError("tied embeddings are not supported with pipeline parallelism; set "
                             "model.share_embeddings_and_output_weights=False (reference known_issues.rst:14-24)")
        skip = {"CoreAttention"} if g.use_flash_attention and g.attention_dropout == 0.0 else set()
        return initialize_parallel_model(
            self, lambda **kw: GPTModel(g, device=kw.pop("device", None), **kw), transformer_layer_cls=ParallelTransformerLayer,
            num_layers=g.num_layers, activation_checkpoint=self._activation_checkpoint_spec(), checkpoint_skip=skip,
            pipeline_cuts=self.config.get("pipeline_cuts", None), hidden_size=g.hidden_size)

    def model_fwd_calc_loss(self, batch: Dict[str, torch.Tensor]):
        ids = batch.get("tokens", batch.get("input_ids"))
        labels = batch["labels"]
        if "tokens" not in batch and self._hf_style_batches():
            # HF-style (SFT / alignment) data modules hand out labels aligned with the inputs and -100 on masked
            # positions; the Megatron model scores next-token labels, so shift here
            labels = torch.cat([labels[:, 1:], labels.new_full((labels.shape[0], 1), -100)], dim=1)
        out = self.model(input_ids=ids, labels=labels, position_ids=batch.get("position_ids"),
                         loss_mask=batch.get("loss_mask"))
        return out[0], out[1]

    def _hf_style_batches(self) -> bool:
        d = self.cfg.get("data", Config({})) if "data" in self.cfg else Config({})
        return self.cfg.get("model_alignment_strategy", None) is not None or bool(d.get("use_sft_style_data_module", False))

    def to_device(self, batch):
        b = super().to_device(batch)
        if "tokens" in b and "input_ids" not in b:
            b["input_ids"] = b["tokens"]
        return b
