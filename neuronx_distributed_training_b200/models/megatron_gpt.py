"""Megatron-style GPT family (GPT-2 style, Llama-2, Mistral w/ sliding window, Mixtral MoE).

Capability parity with the reference's NeMo-derived model (reference
src/.../models/megatron/{gpt_model,language_model,transformer,rotary_pos_embedding}.py):
vocab-parallel embedding (+ learned absolute positions), ``[b,s,h] -> [s,b,h]`` with sequence-parallel scatter and
dropout under the TP RNG fork; fused-QKV or Q + KV (GQA/MQA) attention; RoPE with rotary_percentage and the
position-interpolation / ABF factors; gelu / geglu / reglu / swiglu MLP with optional biases; MoE MLP
(top-k or sinkhorn router); pre_ln / post_ln / normformer / gpt_j block types; LayerNorm / LayerNorm1P / RMSNorm;
tied or untied output layer; loss = masked mean of vocab-parallel CE (+ MoE aux loss).
Parameter names follow the Megatron checkpoint convention so the converters keep working
(``language_model.encoder.layers.N.self_attention.query_key_value`` ...).

Unlike the reference, the Megatron models DO use the flash-attention kernel (the reference documents the lack
of it as a known issue, docs/general/known_issues.rst:161-162); attention dropout / query-key layer scaling fall
back to the eager path that supports them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..parallel import state as ps
from ..parallel import mappings as mp
from ..parallel.layers import ColumnParallelLinear, RowParallelLinear, ParallelEmbedding
from ..parallel.loss import parallel_cross_entropy, fused_linear_cross_entropy
from ..parallel.norm import get_norm
from ..parallel.random import get_rng_tracker


@dataclass
class GPTConfig:
    vocab_size: int = 50304
    hidden_size: int = 4096
    ffn_hidden_size: Optional[int] = None
    num_layers: int = 32
    num_attention_heads: int = 32
    num_kv_heads: Optional[int] = None            # None => MHA
    kv_channels: Optional[int] = None
    max_position_embeddings: int = 2048
    encoder_seq_length: int = 2048
    init_method_std: float = 0.02
    use_scaled_init_method: bool = True
    hidden_dropout: float = 0.1
    attention_dropout: float = 0.1
    ffn_dropout: float = 0.0
    apply_query_key_layer_scaling: bool = False
    normalization: str = "layernorm"             # layernorm | layernorm1p | rmsnorm
    layernorm_epsilon: float = 1e-5
    persist_layer_norm: bool = True
    share_embeddings_and_output_weights: bool = True
    position_embedding_type: str = "learned_absolute"   # learned_absolute | rope
    rotary_percentage: float = 1.0
    rotary_base: float = 10000.0
    position_interpolation_factor: float = 1.0
    position_abf_factor: float = 1.0
    activation: str = "gelu"
    transformer_block_type: str = "pre_ln"       # pre_ln | post_ln | normformer | gpt_j
    has_bias: bool = True
    bias: Optional[bool] = None
    sliding_window_size: Optional[int] = None
    sequence_parallel: bool = False
    use_flash_attention: bool = True
    fused_linear_ce: bool = True
    params_dtype: torch.dtype = torch.bfloat16
    use_cpu_initialization: bool = False
    # MoE
    num_moe_experts: int = 1
    moe_frequency: int = 1
    moe_top_k: int = 2
    moe_routing_algorithm: str = "top_k"
    moe_router_activation: str = "softmax"
    moe_capacity_factor: float = 0.0
    moe_dropless: bool = True
    moe_sinkhorn_iterations: int = 30
    moe_sinkhorn_tol: Optional[float] = None
    moe_aux_loss_coef: float = 0.0
    moe_normalize_top_k_affinities: bool = True
    moe_glu_mlp: bool = True
    expert_model_parallel_size: int = 1
    token_shuffle_group_size: int = 1

    def __post_init__(self):
        if self.ffn_hidden_size is None:
            self.ffn_hidden_size = 4 * self.hidden_size
        if self.kv_channels is None:
            self.kv_channels = self.hidden_size // self.num_attention_heads
        if self.num_kv_heads is None:
            self.num_kv_heads = self.num_attention_heads
        if self.bias is not None:
            self.has_bias = bool(self.bias)

    @property
    def is_glu(self) -> bool:
        return self.activation in ("swiglu", "geglu", "reglu", "fast-swiglu", "fast-geglu", "fast-reglu")


def init_method_normal(std: float):
    return lambda t: nn.init.normal_(t, 0.0, std)


def scaled_init_method_normal(std: float, num_layers: int):
    s = std / math.sqrt(2.0 * num_layers)
    return lambda t: nn.init.normal_(t, 0.0, s)


class RotaryEmbedding(nn.Module):
    """rotary_percentage, position interpolation and ABF (base scaling) as in the reference's
    rotary_pos_embedding.py:29-56."""

    def __init__(self, dim: int, base: float = 10000.0, interpolation_factor: float = 1.0, abf_factor: float = 1.0):
        super().__init__()
        self.dim = dim
        inv_freq = 1.0 / ((base * abf_factor) ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer("inv_freq", inv_freq, persistent=False)
        self.interp = interpolation_factor

    @torch.no_grad()
    def forward(self, seq_len: int, offset: int = 0, device=None):
        t = (torch.arange(offset, offset + seq_len, device=device or self.inv_freq.device).float()) / self.interp
        freqs = torch.outer(t, self.inv_freq.to(t.device))
        emb = torch.cat((freqs, freqs), dim=-1)
        return emb.cos(), emb.sin()


def apply_partial_rope(q, k, cos, sin, rot_dim: int):
    """q/k [B,S,H,D]; rotate the first ``rot_dim`` channels."""
    D = q.shape[-1]
    if rot_dim == D:
        return ops.rope(q, k, cos, sin)
    q1, q2 = q[..., :rot_dim], q[..., rot_dim:]
    k1, k2 = k[..., :rot_dim], k[..., rot_dim:]
    q1, k1 = ops.rope_ref(q1, cos, sin), ops.rope_ref(k1, cos, sin)
    return torch.cat([q1, q2], -1), torch.cat([k1, k2], -1)


class Embedding(nn.Module):
    def __init__(self, cfg: GPTConfig, device=None):
        super().__init__()
        self.cfg = cfg
        init = init_method_normal(cfg.init_method_std)
        self.word_embeddings = ParallelEmbedding(cfg.vocab_size, cfg.hidden_size, init_method=init, dtype=cfg.params_dtype,
                                                 device=device, sequence_parallel_enabled=False,
                                                 use_cpu_initialization=cfg.use_cpu_initialization)
        if cfg.position_embedding_type == "learned_absolute":
            self.position_embeddings = nn.Embedding(cfg.max_position_embeddings, cfg.hidden_size, dtype=cfg.params_dtype, device=device)
            init(self.position_embeddings.weight)
        else:
            self.position_embeddings = None
        self.dropout = nn.Dropout(cfg.hidden_dropout)

    def forward(self, input_ids: torch.Tensor, position_ids: Optional[torch.Tensor]) -> torch.Tensor:
        h = self.word_embeddings(input_ids)                         # [b,s,h]
        if self.position_embeddings is not None:
            if position_ids is None:
                position_ids = torch.arange(input_ids.shape[1], device=input_ids.device).unsqueeze(0)
            h = h + self.position_embeddings(position_ids)
        h = h.transpose(0, 1).contiguous()                          # [s,b,h]
        if self.cfg.sequence_parallel:
            h = mp.scatter_to_sequence_parallel_region(h)
            with get_rng_tracker().fork():
                h = self.dropout(h)
        else:
            h = self.dropout(h)
        return h


class ParallelMLP(nn.Module):
    def __init__(self, cfg: GPTConfig, layer_number: int, device=None):
        super().__init__()
        self.cfg = cfg
        init = init_method_normal(cfg.init_method_std)
        out_init = scaled_init_method_normal(cfg.init_method_std, cfg.num_layers) if cfg.use_scaled_init_method else init
        mult = 2 if cfg.is_glu else 1
        kw = dict(dtype=cfg.params_dtype, device=device, sequence_parallel_enabled=cfg.sequence_parallel,
                  use_cpu_initialization=cfg.use_cpu_initialization)
        self.dense_h_to_4h = ColumnParallelLinear(cfg.hidden_size, mult * cfg.ffn_hidden_size, bias=cfg.has_bias,
                                                  gather_output=False, init_method=init, stride=mult, skip_bias_add=True, **kw)
        self.dense_4h_to_h = RowParallelLinear(cfg.ffn_hidden_size, cfg.hidden_size, bias=cfg.has_bias, input_is_parallel=True,
                                               init_method=out_init, skip_bias_add=True, **kw)
        self.dropout = nn.Dropout(cfg.ffn_dropout) if cfg.ffn_dropout > 0 else None

    def forward(self, x):
        h, b = self.dense_h_to_4h(x)
        h = ops.bias_act(h, b, self.cfg.activation)
        if self.dropout is not None:
            h = self.dropout(h)
        return self.dense_4h_to_h(h)       # (out, bias)


class CoreAttention(nn.Module):
    """MHA / MQA / GQA core; causal or sliding-window mask; query-key layer scaling; attention dropout under
    the TP RNG fork (reference transformer.py:470-777)."""

    def __init__(self, cfg: GPTConfig, layer_number: int):
        super().__init__()
        self.cfg = cfg
        self.layer_number = max(1, layer_number)
        self.scale = 1.0 / math.sqrt(cfg.kv_channels)
        self.dropout_p = cfg.attention_dropout
        self.window = cfg.sliding_window_size

    def forward(self, q, k, v):
        cfg = self.cfg
        needs_eager = (self.dropout_p > 0.0 and self.training) or not cfg.use_flash_attention
        if not needs_eager:
            # layer scaling cancels analytically (scores / n then softmax in fp32 * n), so the flash path is exact
            return ops.flash_attention(q, k, v, causal=True, window=self.window, scale=self.scale)
        with get_rng_tracker().fork():
            if q.is_cuda:
                # scores in the compute dtype + fused scale/mask/softmax kernel (apex FusedScaleMaskSoftmax role)
                return ops.attention_eager(q, k, v, causal=True, window=self.window, scale=self.scale,
                                           dropout_p=self.dropout_p if self.training else 0.0)
            return ops.attention_ref(q, k, v, causal=True, window=self.window, scale=self.scale,
                                     dropout_p=self.dropout_p if self.training else 0.0)


class ParallelAttention(nn.Module):
    def __init__(self, cfg: GPTConfig, layer_number: int, device=None):
        super().__init__()
        self.cfg = cfg
        tp = ps.get_tensor_model_parallel_size()
        self.hd = cfg.kv_channels
        self.nh, self.nkv = cfg.num_attention_heads, cfg.num_kv_heads
        assert self.nh % tp == 0
        self.nh_local = self.nh // tp
        init = init_method_normal(cfg.init_method_std)
        out_init = scaled_init_method_normal(cfg.init_method_std, cfg.num_layers) if cfg.use_scaled_init_method else init
        kw = dict(dtype=cfg.params_dtype, device=device, sequence_parallel_enabled=cfg.sequence_parallel,
                  use_cpu_initialization=cfg.use_cpu_initialization)
        self.gqa = self.nkv != self.nh
        if not self.gqa:
            # fused [q|k|v] interleaved per head: weight rows ordered (head, 3, hd)
            self.query_key_value = ColumnParallelLinear(cfg.hidden_size, 3 * self.nh * self.hd, bias=cfg.has_bias,
                                                        gather_output=False, init_method=init, **kw)
            self.nkv_local = self.nh_local
        else:
            assert self.nkv % tp == 0, "Megatron GQA path needs num_kv_heads % tp == 0"
            self.nkv_local = self.nkv // tp
            self.query = ColumnParallelLinear(cfg.hidden_size, self.nh * self.hd, bias=cfg.has_bias, gather_output=False,
                                              init_method=init, **kw)
            self.key_value = ColumnParallelLinear(cfg.hidden_size, 2 * self.nkv * self.hd, bias=cfg.has_bias,
                                                  gather_output=False, init_method=init, **kw)
        self.core_attention = CoreAttention(cfg, layer_number)
        self.dense = RowParallelLinear(self.nh * self.hd, cfg.hidden_size, bias=cfg.has_bias, input_is_parallel=True,
                                       init_method=out_init, skip_bias_add=True, **kw)
        self.rot_dim = int(self.hd * cfg.rotary_percentage) if cfg.position_embedding_type == "rope" else 0

    def forward(self, hidden, rotary: Optional[Tuple[torch.Tensor, torch.Tensor]]):
        if not self.gqa:
            qkv = self.query_key_value(hidden)                                  # [s,b,nh_local*3*hd]
            S, B = qkv.shape[:2]
            qkv = qkv.view(S, B, self.nh_local, 3 * self.hd)
            q, k, v = qkv.split(self.hd, dim=-1)
        else:
            q = self.query(hidden)
            S, B = q.shape[:2]
            q = q.view(S, B, self.nh_local, self.hd)
            kv = self.key_value(hidden).view(S, B, self.nkv_local, 2 * self.hd)
            k, v = kv.split(self.hd, dim=-1)
        q, k, v = (t.transpose(0, 1) for t in (q, k, v))                        # [b,s,h,d]
        if rotary is not None:
            q, k = apply_partial_rope(q.contiguous(), k.contiguous(), rotary[0], rotary[1], self.rot_dim)
        o = self.core_attention(q, k, v)
        o = o.transpose(0, 1).reshape(S, B, self.nh_local * self.hd)
        return self.dense(o)                                                     # (out, bias)


class ParallelTransformerLayer(nn.Module):
    def __init__(self, cfg: GPTConfig, layer_number: int, device=None):
        super().__init__()
        self.cfg = cfg
        self.layer_number = layer_number
        self.block = cfg.transformer_block_type
        nk = dict(eps=cfg.layernorm_epsilon, sequence_parallel_enabled=cfg.sequence_parallel, dtype=cfg.params_dtype, device=device)
        self.input_layernorm = get_norm(cfg.normalization, cfg.hidden_size, **nk)
        self.self_attention = ParallelAttention(cfg, layer_number, device)
        if self.block == "normformer":
            self.post_attention_normformer_norm = get_norm(cfg.normalization, cfg.hidden_size, **nk)
        if self.block != "gpt_j":
            self.post_attention_layernorm = get_norm(cfg.normalization, cfg.hidden_size, **nk)
        is_moe = cfg.num_moe_experts > 1 and (layer_number % max(1, cfg.moe_frequency) == 0)
        if is_moe:
            from .moe import NeuronSwitchMLP
            self.mlp = NeuronSwitchMLP(cfg, device)
        else:
            self.mlp = ParallelMLP(cfg, layer_number, device)
        self.is_moe = is_moe
        if self.block == "post_ln":
            self.final_layernorm_post = None
        self.hidden_dropout = cfg.hidden_dropout

    def _bda(self, x, bias, residual):
        if self.cfg.sequence_parallel and self.hidden_dropout > 0 and self.training:
            with get_rng_tracker().fork():
                return ops.bias_dropout_add(x, bias, residual, self.hidden_dropout, True)
        return ops.bias_dropout_add(x, bias, residual, self.hidden_dropout, self.training)

    def forward(self, hidden, rotary):
        aux = None
        if self.block == "gpt_j":
            ln = self.input_layernorm(hidden)
            a, ab = self.self_attention(ln, rotary)
            m = self.mlp(ln)
            m, mb = m if not self.is_moe else (m[0], None)
            if self.is_moe:
                aux = self.mlp.last_aux_loss
            out = self._bda(a, ab, hidden)
            return self._bda(m, mb, out), aux
        if self.block == "post_ln":
            a, ab = self.self_attention(hidden, rotary)
            h = self.input_layernorm(self._bda(a, ab, hidden))
            m = self.mlp(h)
            m, mb = m if not self.is_moe else (m[0], None)
            if self.is_moe:
                aux = self.mlp.last_aux_loss
            return self.post_attention_layernorm(self._bda(m, mb, h)), aux
        # pre_ln / normformer
        a, ab = self.self_attention(self.input_layernorm(hidden), rotary)
        if self.block == "normformer":
            a = self.post_attention_normformer_norm(a if ab is None else a + ab)
            ab = None
        h = self._bda(a, ab, hidden)
        m = self.mlp(self.post_attention_layernorm(h))
        m, mb = m if not self.is_moe else (m[0], None)
        if self.is_moe:
            aux = self.mlp.last_aux_loss
        return self._bda(m, mb, h), aux


class ParallelTransformer(nn.Module):
    def __init__(self, cfg: GPTConfig, device=None, layer_range: Optional[Tuple[int, int]] = None, final_norm: bool = True):
        super().__init__()
        self.cfg = cfg
        lo, hi = layer_range if layer_range is not None else (0, cfg.num_layers)
        self.layers = nn.ModuleDict({str(i): ParallelTransformerLayer(cfg, i + 1, device) for i in range(lo, hi)})
        if final_norm and cfg.transformer_block_type != "post_ln":
            self.final_layernorm = get_norm(cfg.normalization, cfg.hidden_size, eps=cfg.layernorm_epsilon,
                                            sequence_parallel_enabled=cfg.sequence_parallel, dtype=cfg.params_dtype, device=device)
        else:
            self.final_layernorm = None

    def forward(self, hidden, rotary):
        aux_total = None
        for layer in self.layers.values():
            hidden, aux = layer(hidden, rotary)
            if aux is not None:
                aux_total = aux if aux_total is None else aux_total + aux
        if self.final_layernorm is not None:
            hidden = self.final_layernorm(hidden)
        return hidden, aux_total


class TransformerLanguageModel(nn.Module):
    def __init__(self, cfg: GPTConfig, device=None, layer_range=None, pre_process: bool = True, post_process: bool = True):
        super().__init__()
        self.cfg = cfg
        self.pre_process, self.post_process = pre_process, post_process
        if pre_process:
            self.embedding = Embedding(cfg, device)
        self.encoder = ParallelTransformer(cfg, device, layer_range, final_norm=post_process)
        if cfg.position_embedding_type == "rope":
            self.rotary_pos_emb = RotaryEmbedding(int(cfg.kv_channels * cfg.rotary_percentage), cfg.rotary_base,
                                                  cfg.position_interpolation_factor, cfg.position_abf_factor)
        else:
            self.rotary_pos_emb = None
        if post_process and not cfg.share_embeddings_and_output_weights:
            self.output_layer = ColumnParallelLinear(cfg.hidden_size, cfg.vocab_size, bias=False, gather_output=False,
                                                     init_method=init_method_normal(cfg.init_method_std),
                                                     dtype=cfg.params_dtype, device=device,
                                                     sequence_parallel_enabled=cfg.sequence_parallel,
                                                     use_cpu_initialization=cfg.use_cpu_initialization)
        else:
            self.output_layer = None

    def forward(self, input_ids=None, position_ids=None, hidden=None, seq_len=None):
        if hidden is None:
            seq_len = input_ids.shape[1]
            hidden = self.embedding(input_ids, position_ids)
        rotary = self.rotary_pos_emb(seq_len, device=hidden.device) if self.rotary_pos_emb is not None else None
        return self.encoder(hidden, rotary)


class GPTModel(nn.Module):
    """forward(input_ids, position_ids, labels, loss_mask) -> (loss, logits|None)."""

    def __init__(self, cfg: GPTConfig, device=None, layer_range=None, with_embedding: bool = True, with_head: bool = True):
        super().__init__()
        self.cfg = cfg
        self.language_model = TransformerLanguageModel(cfg, device, layer_range, with_embedding, with_head)
        self.with_head = with_head
        if with_head and self.language_model.output_layer is None and hasattr(self.language_model, "embedding"):
            # tied output layer: the embedding weight receives two gradient contributions per backward
            self.language_model.embedding.word_embeddings.weight._multi_use = True

    @property
    def lm_head(self):
        """the pipeline wrapper probes ``.lm_head`` to find the loss-computing stage"""
        return True if self.with_head else None

    def output_weight(self) -> torch.Tensor:
        lm = self.language_model
        if lm.output_layer is not None:
            return lm.output_layer.weight
        if not hasattr(lm, "embedding"):
            raise RuntimeError("tied embeddings are not supported across pipeline stages "
                               "(set share_embeddings_and_output_weights: False); reference known_issues.rst:14-24")
        w = lm.embedding.word_embeddings.weight
        w._multi_use = True                    # tied: written by the embedding backward AND the output-layer backward
        return w

    def forward(self, input_ids=None, labels=None, position_ids=None, loss_mask=None, attention_mask=None,
                hidden=None, seq_len=None, return_logits: bool = False):
        hidden, aux = self.language_model(input_ids, position_ids, hidden, seq_len)
        if not self.with_head:
            return hidden
        cfg = self.cfg
        w = self.output_weight()
        group = ps.get_tensor_model_parallel_group()
        if cfg.sequence_parallel:
            h = mp.gather_from_sequence_parallel_region(hidden, True, group)
        else:
            h = mp.copy_to_tensor_model_parallel_region(hidden, group)
        if labels is None:
            logits = F.linear(h, w)
            return None, logits.transpose(0, 1)
        tgt = labels.transpose(0, 1).contiguous()            # Megatron datasets hand out next-token labels already
        if loss_mask is not None:
            tgt = tgt.masked_fill(loss_mask.transpose(0, 1) <= 0, -100)
        if cfg.fused_linear_ce and not return_logits:
            loss_sum, n = fused_linear_cross_entropy(h.reshape(-1, h.shape[-1]), w, tgt.reshape(-1), group)
            loss = loss_sum / n.clamp(min=1).to(loss_sum.dtype)
            logits = None
        else:
            logits = F.linear(h, w)
            per_tok = parallel_cross_entropy(logits.float(), tgt, group)
            loss = per_tok.sum() / (tgt != -100).sum().clamp(min=1)
            logits = logits.transpose(0, 1) if return_logits else None
        if aux is not None and cfg.moe_aux_loss_coef > 0:
            loss = loss + cfg.moe_aux_loss_coef * aux
        return loss, logits
