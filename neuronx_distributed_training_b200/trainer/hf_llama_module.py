"""HF-Llama task module (pre-training / SFT / LoRA / DPO / ORPO share it).

Capability parity: reference hf_models/base_model.py:18-54 and hf_models/llama_model.py:25-144
(config.json + YAML overrides, SP / qkv_linear / fuse_qkv / kv_replicator / flash / ring flags, LoRA
config, reduce dtype, activation-recompute class list resolved by class *name*, pipeline config).
"""
from __future__ import annotations

import logging
import os
from typing import Optional

import torch
import torch.nn as nn

from ..config import Config
from ..models import llama as llama_mod
from ..models.llama import LlamaConfig, LlamaForCausalLM, LlamaDecoderLayer
from ..parallel.model import initialize_parallel_model
from .module import BaseModelModule, get_param_groups_by_weight_decay

log = logging.getLogger("nxdt")
_CFG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "models", "configs")


def resolve_model_config_path(path: Optional[str], model_cfg: Config, family: str = "llama") -> str:
    """Accept a real ``config.json`` path, the name of a built-in config, or infer from the YAML dims."""
    if path and os.path.exists(path):
        return path
    if path:
        cand = os.path.join(_CFG_DIR, os.path.basename(path))
        if os.path.exists(cand):
            return cand
        cand = os.path.join(_CFG_DIR, str(path) + ".json")
        if os.path.exists(cand):
            return cand
    if family == "mixtral":
        return os.path.join(_CFG_DIR, "mixtral_8x7b.json")
    hs, nl = int(model_cfg.get("hidden_size", 4096)), int(model_cfg.get("num_layers", 32))
    theta = float(model_cfg.get("rope_theta", 10000.0))
    maxpos = int(model_cfg.get("max_position_embeddings", 4096))
    table = {(4096, True): "llama3_8b", (8192, True): "llama3_70b", (16384, True): "llama3.1_405b",
             (4096, False): "llama2_7b", (8192, False): "llama2_70b", (768, False): "llama_125m"}
    name = table.get((hs, theta >= 1e5), "llama2_7b")
    if maxpos > 8192 and name in ("llama3_8b", "llama3_70b"):
        name = name.replace("llama3_", "llama3.1_")
    log.info("model.model_config %r not found; using built-in %s.json", path, name)
    return os.path.join(_CFG_DIR, name + ".json")


class BaseHfModel(BaseModelModule):
    """Shared HF behaviour: weight-decay groups, DPO / ORPO dispatch."""

    def setup_optimizer_param_groups(self):
        wd = float(self.config.optim.get("weight_decay", 0.01))
        # Norm weights are excluded from decay unless the user asks to decay them.  (The reference
        # inverts this flag by accident - SURVEY Appendix B - we follow the flag's name.)
        no_decay_norm = not bool(self.config.get("do_layer_norm_weight_decay", False))
        return get_param_groups_by_weight_decay(self.local_named_parameters(), wd, no_decay_norm)

    def _alignment(self):
        return self.cfg.get("model_alignment_strategy", None)

    def on_train_start(self, trainer):
        a = self._alignment()
        if a is not None and a.get("dpo", None) is not None:
            from .dpo import DPOBaseModel
            DPOBaseModel.on_train_start(self, trainer)

    def model_fwd_calc_loss(self, batch):
        a = self._alignment()
        if a is not None and a.get("dpo", None) is not None:
            from .dpo import DPOBaseModel
            return DPOBaseModel.model_fwd_calc_loss(self, batch)
        if a is not None and a.get("orpo", None) is not None:
            from .orpo import ORPOBaseModel
            return ORPOBaseModel.model_fwd_calc_loss(self, batch)
        return super().model_fwd_calc_loss(batch)


class HFLlamaModule(BaseHfModel):
    model_family = "llama"

    def _get_model_config(self) -> LlamaConfig:
        m = self.config
        ds = self.cfg.get("distributed_strategy", Config({}))
        path = resolve_model_config_path(m.get("model_config", None), m, self.model_family)
        hf = LlamaConfig.from_pretrained(path)
        for yaml_key, attr in (("num_layers", "num_hidden_layers"), ("hidden_size", "hidden_size"),
                               ("rope_theta", "rope_theta"), ("max_position_embeddings", "max_position_embeddings"),
                               ("head_dim", "head_dim"), ("vocab_size", "vocab_size"),
                               ("intermediate_size", "intermediate_size"), ("ffn_hidden_size", "intermediate_size"),
                               ("num_attention_heads", "num_attention_heads"), ("num_kv_heads", "num_key_value_heads"),
                               ("sliding_window", "sliding_window"), ("tie_word_embeddings", "tie_word_embeddings")):
            if m.get(yaml_key, None) is not None:
                setattr(hf, attr, type(getattr(hf, attr))(m[yaml_key]) if getattr(hf, attr) is not None else m[yaml_key])
        if m.get("hidden_size", None) is not None and m.get("head_dim", None) is None:
            hf.head_dim = hf.hidden_size // hf.num_attention_heads
        if bool(getattr(hf, "tie_word_embeddings", False)) and self.dims.pp > 1:
            # embedding and LM head land on different stages: the reference fails with "shared weight has None grad"
            # (docs/general/known_issues.rst); reject the combination up front instead
            raise ValueError("tie_word_embeddings is not supported with pipeline parallelism; set "
                             "model.tie_word_embeddings: false (the head then gets its own weight) or use PP=1")
        fus = m.get("fusions", Config({})) or Config({})
        hf.sequence_parallel_enabled = self.sequence_parallel
        hf.qkv_linear = bool(m.get("qkv_linear", True))
        hf.fuse_qkv = bool(m.get("fuse_qkv", True))
        hf.kv_shared_group_size = int(ds.get("kv_replicator", 1) or 1)
        hf.use_flash_attention = bool(fus.get("flash_attention", True))
        hf.use_ring_attention = bool(fus.get("ring_attention", False))
        hf.use_ulysses_attention = bool(fus.get("ulysses_attention", False))
        hf.fused_linear_ce = bool(fus.get("linear_cross_entropy", True)) and not self.save_logits
        hf.ce_chunk_tokens = int(m.get("ce_chunk_tokens", 2048))
        hf.use_cpu_initialization = bool(m.get("use_cpu_initialization", False)) and not torch.cuda.is_available() \
            or bool(m.get("force_cpu_initialization", False))
        hf.torch_dtype = {torch.bfloat16: "bfloat16", torch.float32: "float32"}[self.precision.param_dtype]
        # vocab must divide TP
        tp = self.dims.tp
        if hf.vocab_size % tp != 0:
            hf.vocab_size = (hf.vocab_size + tp - 1) // tp * tp
        return hf

    def _activation_checkpoint_spec(self, hf: LlamaConfig):
        m = self.config
        gran = m.get("activations_checkpoint_granularity", None)
        if gran == "full":
            return "full"
        if gran == "selective":
            names = list(m.get("activations_checkpoint_recompute", None) or [])
            for n in names:
                if not hasattr(llama_mod, n):
                    raise ValueError(f"activations_checkpoint_recompute: unknown module class {n!r}")
            return names
        if gran is None and not hf.use_flash_attention:
            return ["CoreAttention"]          # reference default (llama_model.py:88-89)
        return None

    def model_provider_func(self, hf: LlamaConfig, **stage_kw) -> nn.Module:
        return LlamaForCausalLM(hf, device=stage_kw.pop("device", None), **stage_kw)

    def build_model(self) -> nn.Module:
        hf = self._get_model_config()
        self.hf_config = hf
        spec = self._activation_checkpoint_spec(hf)
        # flash attention never materialises S x S, so "recompute CoreAttention" is already what its
        # backward does; wrapping it again would only re-run the forward kernel.
        skip = {"CoreAttention"} if hf.use_flash_attention else set()
        lora = None
        a = self.cfg.get("model_alignment_strategy", None)
        if a is not None and a.get("peft", None) is not None:
            from ..models.lora import LoraConfig
            lora = LoraConfig.from_config(a.peft)
        model = initialize_parallel_model(
            self, lambda **kw: self.model_provider_func(hf, **kw), transformer_layer_cls=LlamaDecoderLayer,
            num_layers=hf.num_hidden_layers, activation_checkpoint=spec, checkpoint_skip=skip,
            pipeline_cuts=self.config.get("pipeline_cuts", None), lora_config=lora, hidden_size=hf.hidden_size)
        return model

    def init_weights(self, module, device):
        if isinstance(module, llama_mod.LlamaRMSNorm):
            with torch.no_grad():
                module.weight.fill_(1.0)
            return
        super().init_weights(module, device)
