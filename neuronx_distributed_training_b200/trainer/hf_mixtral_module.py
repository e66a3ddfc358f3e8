"""HF-Mixtral task module (reference hf_models/mixtral_model.py:24-96)."""
from __future__ import annotations


from ..config import Config
from ..models import mixtral as mixtral_mod
from ..models.mixtral import MixtralConfig, MixtralForCausalLM, MixtralDecoderLayer
from ..parallel.model import initialize_parallel_model
from .hf_llama_module import HFLlamaModule, resolve_model_config_path


class HFMixtralModule(HFLlamaModule):
    model_family = "mixtral"

    def _get_model_config(self) -> MixtralConfig:
        base = super()._get_model_config()
        m = self.config
        path = resolve_model_config_path(m.get("model_config", None), m, "mixtral")
        hf = MixtralConfig.from_pretrained(path)
        for k, v in base.to_dict().items():          # framework knobs + YAML overrides resolved by the Llama path
            if k in ("num_local_experts", "num_experts_per_tok", "router_aux_loss_coef", "model_type", "sliding_window"):
                continue
            setattr(hf, k, v)
        moe = m.get("moe", Config({})) or Config({})
        hf.num_local_experts = int(moe.get("num_experts", hf.num_local_experts))
        hf.num_experts_per_tok = int(moe.get("top_k", hf.num_experts_per_tok))
        hf.moe_frequency = int(moe.get("frequency", 1))
        hf.capacity_factor = 0.0 if bool(moe.get("dropless", False)) else float(moe.get("capacity_factor", 0.0) or 0.0)
        hf.routing_algorithm = str(moe.get("routing_algorithm", "top_k"))
        hf.router_activation = str(moe.get("router_activation", "softmax"))
        hf.normalize_top_k_affinities = bool(moe.get("normalize_top_k_affinities", True))
        hf.glu_mlp = bool(moe.get("glu_mlp", True))
        hf.router_aux_loss_coef = float(moe.get("router_aux_loss_coef", hf.router_aux_loss_coef))
        hf.output_router_logits = bool(moe.get("output_router_logits", True))
        hf.sinkhorn_iterations = int(moe.get("sinkhorn_iterations", 30))
        hf.sinkhorn_tol = moe.get("sinkhorn_tol", None)
        ds = self.cfg.get("distributed_strategy", Config({}))
        hf.token_shuffle_group_size = int(ds.get("token_shuffle_group_size", 1) or 1)
        if m.get("sliding_window", None) is not None:
            hf.sliding_window = int(m.sliding_window)
        return hf

    def _activation_checkpoint_spec(self, hf):
        m = self.config
        gran = m.get("activations_checkpoint_granularity", None)
        if gran == "selective":
            names = list(m.get("activations_checkpoint_recompute", None) or [])
            for n in names:
                if not (hasattr(mixtral_mod, n) or hasattr(__import__("neuronx_distributed_training_b200.models.llama", fromlist=["x"]), n)):
                    raise ValueError(f"activations_checkpoint_recompute: unknown module class {n!r}")
            return names
        return super()._activation_checkpoint_spec(hf)

    def model_provider_func(self, hf, **stage_kw):
        return MixtralForCausalLM(hf, device=stage_kw.pop("device", None), **stage_kw)

    def build_model(self):
        hf = self._get_model_config()
        self.hf_config = hf
        spec = self._activation_checkpoint_spec(hf)
        skip = {"CoreAttention"} if hf.use_flash_attention else set()
        return initialize_parallel_model(self, lambda **kw: self.model_provider_func(hf, **kw),
                                         transformer_layer_cls=MixtralDecoderLayer, num_layers=hf.num_hidden_layers,
                                         activation_checkpoint=spec, checkpoint_skip=skip,
                                         pipeline_cuts=self.config.get("pipeline_cuts", None), hidden_size=hf.hidden_size)
