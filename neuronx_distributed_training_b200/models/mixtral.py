"""HF-style Mixtral (sparse MoE) built on the Llama blocks.

Capability parity with the reference's ``modeling_mixtral.py`` (reference
src/.../models/hf_models/modeling_mixtral.py:123-154,342-374,451-454,537-547,872-878): sliding-window attention,
``RouterTopK + ExpertMLPs + MoE`` block with ``moe_frequency`` interleave, router logits threaded through the layers
(and across pipeline stages) for the Switch-style load-balancing loss.  Parameter names match the reference converter
(``mlp.moe.router.linear_router``, ``mlp.moe.expert_mlps.mlp_op.{gate_up,down}_proj``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn

from .llama import LlamaConfig, LlamaAttention, LlamaMLP, LlamaRMSNorm, LlamaModel, LlamaForCausalLM
from .moe import RouterTopK, RouterSinkhorn, ExpertMLPs, MoE, load_balancing_loss_func


@dataclass
class MixtralConfig(LlamaConfig):
    num_local_experts: int = 8
    num_experts_per_tok: int = 2
    router_aux_loss_coef: float = 0.02
    output_router_logits: bool = True
    moe_frequency: int = 1
    capacity_factor: float = 0.0
    routing_algorithm: str = "top_k"
    router_activation: str = "softmax"
    normalize_top_k_affinities: bool = True
    glu_mlp: bool = True
    sinkhorn_iterations: int = 30
    sinkhorn_tol: Optional[float] = None
    token_shuffle_group_size: int = 1
    model_type: str = "mixtral"


def initialize_mixtral_moe_layer(config: MixtralConfig, device=None) -> MoE:
    rk = dict(num_experts=config.num_local_experts, top_k=config.num_experts_per_tok, hidden_size=config.hidden_size,
              act_fn=config.router_activation, device=device, dtype=config.dtype)
    router = RouterSinkhorn(**rk, sinkhorn_iterations=config.sinkhorn_iterations, sinkhorn_tol=config.sinkhorn_tol) \
        if config.routing_algorithm == "sinkhorn" else RouterTopK(**rk)
    experts = ExpertMLPs(config.num_local_experts, config.hidden_size, config.intermediate_size, glu=config.glu_mlp,
                         activation="silu", dtype=config.dtype, device=device, init_std=config.initializer_range,
                         use_cpu_initialization=config.use_cpu_initialization)
    return MoE(router, experts, capacity_factor=config.capacity_factor,
               normalize_top_k_affinities=config.normalize_top_k_affinities,
               sequence_parallel_enabled=config.sequence_parallel_enabled, return_router_logits=True,
               token_shuffle_group_size=config.token_shuffle_group_size)


class MixtralSparseMoeBlock(nn.Module):
    def __init__(self, config: MixtralConfig, device=None):
        super().__init__()
        self.moe = initialize_mixtral_moe_layer(config, device)

    def forward(self, x):
        return self.moe(x)        # (out, router_logits)


class MixtralDecoderLayer(nn.Module):
    def __init__(self, config: MixtralConfig, layer_idx: int = 0, device=None):
        super().__init__()
        self.self_attn = LlamaAttention(config, layer_idx, device)
        self.is_moe = (layer_idx % max(1, config.moe_frequency)) == 0
        self.mlp = MixtralSparseMoeBlock(config, device) if self.is_moe else LlamaMLP(config, device)
        sp = config.sequence_parallel_enabled
        self.input_layernorm = LlamaRMSNorm(config.hidden_size, config.rms_norm_eps, sp, config.dtype, device)
        self.post_attention_layernorm = LlamaRMSNorm(config.hidden_size, config.rms_norm_eps, sp, config.dtype, device)

    def forward(self, hidden, cos, sin):
        hidden = hidden + self.self_attn(self.input_layernorm(hidden), cos, sin)
        h = self.post_attention_layernorm(hidden)
        if self.is_moe:
            m, router_logits = self.mlp(h)
        else:
            m, router_logits = self.mlp(h), None
        return hidden + m, router_logits


class MixtralModel(LlamaModel):
    def __init__(self, config: MixtralConfig, device=None, layer_range=None, with_embedding=True, with_final_norm=True):
        super().__init__(config, device, layer_range, with_embedding, with_final_norm)
        lo, hi = self.layer_range
        self.layers = nn.ModuleDict({str(i): MixtralDecoderLayer(config, i, device) for i in range(lo, hi)})

    def forward(self, input_ids=None, hidden=None, seq_len=None):
        if hidden is None:
            seq_len = input_ids.shape[1]
            hidden = self.embed(input_ids)
        cos, sin = self.rotary_emb(self.position_ids(seq_len, hidden.device))
        all_router_logits: List[torch.Tensor] = []
        for layer in self.layers.values():
            hidden, rl = layer(hidden, cos, sin)
            if rl is not None:
                all_router_logits.append(rl)
        if self.norm is not None:
            hidden = self.norm(hidden)
        return hidden, all_router_logits


class MixtralForCausalLM(LlamaForCausalLM):
    def __init__(self, config: MixtralConfig, device=None, layer_range=None, with_embedding=True, with_head=True):
        super().__init__(config, device, layer_range, with_embedding, with_head)
        self.model = MixtralModel(config, device, layer_range, with_embedding, with_final_norm=with_head)
        if with_head and config.tie_word_embeddings and with_embedding:
            self.lm_head.weight = self.model.embed_tokens.weight
        self._pending_aux: Optional[torch.Tensor] = None

    def _aux(self, router_logits: List[torch.Tensor]) -> Optional[torch.Tensor]:
        cfg = self.config
        if not router_logits or cfg.router_aux_loss_coef <= 0:
            return None
        return load_balancing_loss_func(tuple(router_logits), cfg.num_local_experts, cfg.num_experts_per_tok)

    def forward(self, input_ids=None, labels=None, attention_mask=None, hidden=None, seq_len=None,
                return_logits: bool = False, loss_mask=None):
        hidden, router_logits = self.model(input_ids=input_ids, hidden=hidden, seq_len=seq_len)
        aux = self._aux(router_logits)
        if self.lm_head is None:
            # pipeline stage without the head: the local aux loss joins the graph through a zero-valued hook so the
            # router still gets its gradient on this stage (stages exchange only the hidden state)
            if aux is not None and torch.is_grad_enabled():
                hidden = _AttachAux.apply(hidden, aux * (self.config.router_aux_loss_coef /
                                                         getattr(self, "pipeline_num_microbatches", 1)))
            return hidden
        if labels is None:
            return None, self.logits(hidden)
        loss, logits = self.compute_loss(hidden, labels, return_logits, loss_mask)
        if aux is not None:
            loss = loss + self.config.router_aux_loss_coef * aux
        return loss, logits


class _AttachAux(torch.autograd.Function):
    """identity on ``hidden`` whose backward also feeds d(aux)=1/num_microbatches-scaled gradient of the aux loss."""

    @staticmethod
    def forward(ctx, hidden, aux):
        return hidden.view_as(hidden)

    @staticmethod
    def backward(ctx, g):
        return g, torch.ones((), dtype=torch.float32, device=g.device)
