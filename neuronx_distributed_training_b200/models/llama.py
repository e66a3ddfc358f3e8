"""HF-style Llama (2 / 3 / 3.1) built from the tensor-parallel layers.

Capability parity with the reference's ``modeling_llama.py`` (reference
src/.../models/hf_models/modeling_llama.py:145-873): fused gate||up SwiGLU MLP, three QKV layouts
(separate / fused-stride-3 / GQA-QKV with KV replication), RoPE incl. llama3 scaling, eager / flash /
ring attention, sequence parallelism with ``[S/TP, B, H]`` activations, vocab-parallel LM head + CE,
context-parallel position offsets.  Module and parameter names follow HF / the reference so the
checkpoint converters and the selective-recompute class lists keep working.

B200-first differences: q/k/v are produced by ONE GEMM whenever possible; attention runs on
``[B, S, heads, D]`` views with GQA handled inside the kernel (no ``repeat_kv`` materialisation); the
LM head and the loss are fused and chunked so fp32 logits never hit HBM.
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, asdict
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from ..parallel import state as ps
from ..parallel import mappings as mp
from ..parallel.layers import (ColumnParallelLinear, RowParallelLinear, GQAQKVColumnParallelLinear,
                               ParallelEmbedding)
from ..parallel.loss import parallel_cross_entropy, fused_linear_cross_entropy
from ..parallel.norm import RMSNorm


@dataclass
class LlamaConfig:
    vocab_size: int = 32000
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: Optional[int] = None
    head_dim: Optional[int] = None
    hidden_act: str = "silu"
    max_position_embeddings: int = 4096
    initializer_range: float = 0.02
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    rope_scaling: Optional[Dict[str, Any]] = None
    attention_bias: bool = False
    mlp_bias: bool = False
    tie_word_embeddings: bool = False
    pad_token_id: Optional[int] = None
    bos_token_id: int = 1
    eos_token_id: Any = 2
    attention_dropout: float = 0.0
    sliding_window: Optional[int] = None
    # ---- framework knobs (set by the model module from the YAML) --------------------------------
    sequence_parallel_enabled: bool = False
    qkv_linear: bool = True              # use the GQA-QKV module (KV replication aware)
    fuse_qkv: bool = True
    kv_shared_group_size: int = 1        # = distributed_strategy.kv_replicator
    use_flash_attention: bool = True
    use_ring_attention: bool = False
    use_ulysses_attention: bool = False     # all-to-all context parallelism instead of the ring (parallel/ulysses_attention.py)
    fused_linear_ce: bool = True
    ce_chunk_tokens: int = 2048
    torch_dtype: str = "bfloat16"
    use_cpu_initialization: bool = True
    reduce_dtype: Optional[str] = None
    model_type: str = "llama"

    def __post_init__(self):
        if self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads
        if self.head_dim is None:
            self.head_dim = self.hidden_size // self.num_attention_heads

    @property
    def dtype(self) -> torch.dtype:
        return {"bfloat16": torch.bfloat16, "bf16": torch.bfloat16, "float32": torch.float32,
                "fp32": torch.float32, "float16": torch.float16}[str(self.torch_dtype).replace("torch.", "")]

    @classmethod
    def from_dict(cls, d: Dict[str, Any]) -> "LlamaConfig":
        known = {f for f in cls.__dataclass_fields__}
        return cls(**{k: v for k, v in d.items() if k in known})

    @classmethod
    def from_pretrained(cls, path: str, **overrides) -> "LlamaConfig":
        """``path`` is a HF ``config.json`` (or a directory containing one)."""
        import os
        if os.path.isdir(path):
            path = os.path.join(path, "config.json")
        with open(path) as f:
            d = json.load(f)
        d.update(overrides)
        return cls.from_dict(d)

    def to_dict(self) -> Dict[str, Any]:
        return asdict(self)


# ---------------------------------------------------------------------------------------------------
class LlamaRMSNorm(RMSNorm):
    """fp32-math RMSNorm; weight tagged for the SP gradient all-reduce (reference :145-161)."""

    def __init__(self, hidden_size, eps=1e-6, sequence_parallel_enabled=False, dtype=torch.float32, device=None):
        super().__init__(hidden_size, eps, sequence_parallel_enabled, dtype, device)


class LlamaRotaryEmbedding(nn.Module):
    """cos/sin tables in fp32 (reference :847-873), incl. the llama3 / linear / dynamic-NTK scalings."""

    def __init__(self, config: LlamaConfig, device=None):
        super().__init__()
        dim = config.head_dim
        base = float(config.rope_theta)
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.float64) / dim))
        self.attention_scaling = 1.0
        rs = config.rope_scaling
        if rs:
            rtype = rs.get("rope_type", rs.get("type", "default"))
            if rtype == "linear":
                inv_freq = inv_freq / float(rs["factor"])
            elif rtype == "llama3":
                factor, lo, hi = float(rs["factor"]), float(rs.get("low_freq_factor", 1.0)), float(rs.get("high_freq_factor", 4.0))
                old_len = float(rs.get("original_max_position_embeddings", 8192))
                wavelen = 2 * math.pi / inv_freq
                smooth = ((old_len / wavelen) - lo) / (hi - lo)
                scaled = torch.where(wavelen > old_len / lo, inv_freq / factor, inv_freq)
                mid = (wavelen <= old_len / lo) & (wavelen >= old_len / hi)
                scaled = torch.where(mid, (1 - smooth) * inv_freq / factor + smooth * inv_freq, scaled)
                inv_freq = scaled
            elif rtype == "dynamic":
                factor = float(rs["factor"])
                seq = config.max_position_embeddings
                base2 = base * ((factor * seq / config.max_position_embeddings) - (factor - 1)) ** (dim / (dim - 2))
                inv_freq = 1.0 / (base2 ** (torch.arange(0, dim, 2, dtype=torch.float64) / dim))
            elif rtype == "yarn":
                self.attention_scaling = 0.1 * math.log(float(rs["factor"])) + 1.0
                inv_freq = inv_freq / float(rs["factor"])
        self.register_buffer("inv_freq", inv_freq.float().to(device) if device is not None else inv_freq.float(),
                             persistent=False)

    @torch.no_grad()
    def forward(self, position_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """position_ids [S] (shared across the batch) -> cos, sin [S, D] fp32."""
        freqs = torch.outer(position_ids.float(), self.inv_freq.float().to(position_ids.device))
        emb = torch.cat((freqs, freqs), dim=-1)
        return emb.cos() * self.attention_scaling, emb.sin() * self.attention_scaling


class ActivationMultiplyMLP(nn.Module):
    """SwiGLU on the fused gate||up activation (reference :164-173).

    Selecting this class for activation recomputation (``activations_checkpoint_recompute``) does not wrap it
    in a checkpoint: ``enable_recompute`` makes the consuming ``down_proj`` node compute the activation itself
    and recompute it in backward from the saved gate||up tensor."""

    recompute_in_consumer = False

    def enable_recompute(self) -> bool:
        self.recompute_in_consumer = True
        return True

    def forward(self, gate_up: torch.Tensor) -> torch.Tensor:
        return ops.swiglu(gate_up)


class LlamaMLP(nn.Module):
    def __init__(self, config: LlamaConfig, device=None):
        super().__init__()
        std = config.initializer_range
        init = lambda t: nn.init.normal_(t, 0.0, std)
        kw = dict(dtype=config.dtype, device=device, sequence_parallel_enabled=config.sequence_parallel_enabled,
                  use_cpu_initialization=config.use_cpu_initialization)
        self.gate_up_proj = ColumnParallelLinear(config.hidden_size, 2 * config.intermediate_size, bias=config.mlp_bias,
                                                 gather_output=False, init_method=init, stride=2, **kw)
        self.activation_multiply = ActivationMultiplyMLP()
        self.down_proj = RowParallelLinear(config.intermediate_size, config.hidden_size, bias=config.mlp_bias,
                                           input_is_parallel=True, init_method=init, **kw)

    def forward(self, x, pre_norm: Optional[nn.Module] = None):
        # ``pre_norm``: the RMSNorm in front of the block, handed to the column linear so that norm + sequence-parallel
        # all-gather can run as one NVLS multicast kernel (parallel/nvls.py); applied here when the linear cannot take it
        if pre_norm is not None and type(self.gate_up_proj) is not ColumnParallelLinear:
            x, pre_norm = pre_norm(x), None
        gate_up = self.gate_up_proj(x, pre_norm) if pre_norm is not None else self.gate_up_proj(x)
        if self.activation_multiply.recompute_in_consumer and torch.is_grad_enabled() \
                and type(self.down_proj) is RowParallelLinear:      # not when wrapped (e.g. LoRA)
            return self.down_proj(gate_up, pre_act="swiglu")
        return self.down_proj(self.activation_multiply(gate_up))


class CoreAttention(nn.Module):
    """softmax(QK^T/sqrt(d))V on ``[B,S,heads,D]``; flash kernel, eager fallback, sliding window."""

    def __init__(self, config: LlamaConfig):
        super().__init__()
        self.use_flash = config.use_flash_attention
        self.window = config.sliding_window
        self.scale = 1.0 / math.sqrt(config.head_dim)

    def forward(self, q, k, v):
        if self.use_flash:
            return ops.flash_attention(q, k, v, causal=True, window=self.window, scale=self.scale)
        return ops.attention_ref(q, k, v, causal=True, window=self.window, scale=self.scale)


class LlamaAttention(nn.Module):
    def __init__(self, config: LlamaConfig, layer_idx: int = 0, device=None):
        super().__init__()
        self.config = config
        self.layer_idx = layer_idx
        tp = ps.get_tensor_model_parallel_size()
        self.head_dim = config.head_dim
        self.num_heads, self.num_kv_heads = config.num_attention_heads, config.num_key_value_heads
        std = config.initializer_range
        init = lambda t: nn.init.normal_(t, 0.0, std)
        kw = dict(dtype=config.dtype, device=device, sequence_parallel_enabled=config.sequence_parallel_enabled,
                  use_cpu_initialization=config.use_cpu_initialization)
        self.qkv_mode = "gqa" if config.qkv_linear else ("fused" if config.fuse_qkv and
                                                         self.num_heads == self.num_kv_heads else "separate")
        if self.qkv_mode == "gqa":
            self.qkv_proj = GQAQKVColumnParallelLinear(config.hidden_size, self.num_heads, self.num_kv_heads,
                                                       self.head_dim, bias=config.attention_bias,
                                                       kv_size_multiplier=config.kv_shared_group_size,
                                                       fuse_qkv=config.fuse_qkv, init_method=init, **kw)
            self.local_q_heads = self.num_heads // tp
            self.local_kv_heads = self.num_kv_heads * config.kv_shared_group_size // tp
        elif self.qkv_mode == "fused":
            self.qkv_proj = ColumnParallelLinear(config.hidden_size, 3 * self.num_heads * self.head_dim,
                                                 bias=config.attention_bias, gather_output=False,
                                                 init_method=init, stride=3, **kw)
            self.local_q_heads = self.local_kv_heads = self.num_heads // tp
        else:
            assert self.num_kv_heads % tp == 0, "separate q/k/v needs num_kv_heads % tp == 0; use qkv_linear"
            self.q_proj = ColumnParallelLinear(config.hidden_size, self.num_heads * self.head_dim,
                                               bias=config.attention_bias, init_method=init, **kw)
            self.k_proj = ColumnParallelLinear(config.hidden_size, self.num_kv_heads * self.head_dim,
                                               bias=config.attention_bias, init_method=init, **kw)
            self.v_proj = ColumnParallelLinear(config.hidden_size, self.num_kv_heads * self.head_dim,
                                               bias=config.attention_bias, init_method=init, **kw)
            self.local_q_heads, self.local_kv_heads = self.num_heads // tp, self.num_kv_heads // tp
        self.o_proj = RowParallelLinear(self.num_heads * self.head_dim, config.hidden_size, bias=config.attention_bias,
                                        input_is_parallel=True, init_method=init, **kw)
        self.core_attn = CoreAttention(config)
        cp = ps.get_context_model_parallel_size()
        self.use_ulysses = bool(config.use_ulysses_attention) and cp > 1
        self.use_ring = config.use_ring_attention and cp > 1 and not self.use_ulysses
        if cp > 1 and not (config.use_ring_attention or config.use_ulysses_attention):
            raise ValueError("context parallelism requires model.fusions.ring_attention=True "
                             "(reference modeling_llama.py:280-288) or model.fusions.ulysses_attention=True")

    def forward(self, hidden: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor,
                pre_norm: Optional[nn.Module] = None) -> torch.Tensor:
        # hidden: [S_local, B, H]; ``pre_norm``: see LlamaMLP.forward
        if pre_norm is not None and (self.qkv_mode not in ("gqa", "fused") or type(self.qkv_proj).__name__ not in
                                     ("GQAQKVColumnParallelLinear", "ColumnParallelLinear")):
            hidden, pre_norm = pre_norm(hidden), None
        if self.qkv_mode == "gqa" and self.config.fuse_qkv and not self.use_ring and not self.use_ulysses and \
                type(self.qkv_proj).__name__ == "GQAQKVColumnParallelLinear":
            # one GEMM, then one fused kernel: split + RoPE + [S,B]->[B,S] layout, no intermediate copies
            packed = self.qkv_proj.forward_packed(hidden, pre_norm)
            S, B = packed.shape[0], packed.shape[1]
            q, k, v = ops.qkv_rope_split(packed, cos, sin, self.local_q_heads, self.local_kv_heads, self.head_dim)
            o = self.core_attn(q, k, v)
            o = o.transpose(0, 1).reshape(S, B, self.local_q_heads * self.head_dim)
            return self.o_proj(o)
        if self.qkv_mode == "gqa":
            q, k, v = self.qkv_proj(hidden, pre_norm) if pre_norm is not None else self.qkv_proj(hidden)
        elif self.qkv_mode == "fused":
            q, k, v = (self.qkv_proj(hidden, pre_norm) if pre_norm is not None else self.qkv_proj(hidden)).chunk(3, dim=-1)
        else:
            q, k, v = self.q_proj(hidden), self.k_proj(hidden), self.v_proj(hidden)
        S, B = q.shape[0], q.shape[1]
        q = q.reshape(S, B, self.local_q_heads, self.head_dim).transpose(0, 1)
        k = k.reshape(S, B, self.local_kv_heads, self.head_dim).transpose(0, 1)
        v = v.reshape(S, B, self.local_kv_heads, self.head_dim).transpose(0, 1)
        q, k = ops.rope(q, k, cos, sin)
        if self.use_ulysses:
            from ..parallel.ulysses_attention import ulysses_attention
            o = ulysses_attention(q.contiguous(), k.contiguous(), v.contiguous(), self.core_attn)
        elif self.use_ring:
            from ..parallel.ring_attention import ring_attention
            o = ring_attention(q, k, v, scale=self.core_attn.scale)
        else:
            o = self.core_attn(q, k, v)
        o = o.transpose(0, 1).reshape(S, B, self.local_q_heads * self.head_dim)
        return self.o_proj(o)


class LlamaDecoderLayer(nn.Module):
    def __init__(self, config: LlamaConfig, layer_idx: int = 0, device=None):
        super().__init__()
        self.self_attn = LlamaAttention(config, layer_idx, device)
        self.mlp = LlamaMLP(config, device)
        sp = config.sequence_parallel_enabled
        self.input_layernorm = LlamaRMSNorm(config.hidden_size, config.rms_norm_eps, sp, config.dtype, device)
        self.post_attention_layernorm = LlamaRMSNorm(config.hidden_size, config.rms_norm_eps, sp, config.dtype, device)
        self.nvls_pre_norm = False

    def forward(self, hidden: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
        if self.nvls_pre_norm:        # parallel/nvls.py: the norms run inside the column linears (norm + all-gather in one kernel)
            hidden = hidden + self.self_attn(hidden, cos, sin, pre_norm=self.input_layernorm)
            hidden = hidden + self.mlp(hidden, pre_norm=self.post_attention_layernorm)
            return hidden
        hidden = hidden + self.self_attn(self.input_layernorm(hidden), cos, sin)
        hidden = hidden + self.mlp(self.post_attention_layernorm(hidden))
        return hidden


class LlamaModel(nn.Module):
    def __init__(self, config: LlamaConfig, device=None, layer_range: Optional[Tuple[int, int]] = None,
                 with_embedding: bool = True, with_final_norm: bool = True):
        super().__init__()
        self.config = config
        std = config.initializer_range
        init = lambda t: nn.init.normal_(t, 0.0, std)
        lo, hi = layer_range if layer_range is not None else (0, config.num_hidden_layers)
        self.layer_range = (lo, hi)
        if with_embedding:
            self.embed_tokens = ParallelEmbedding(config.vocab_size, config.hidden_size, config.pad_token_id,
                                                  init_method=init, dtype=config.dtype, device=device,
                                                  sequence_parallel_enabled=config.sequence_parallel_enabled,
                                                  use_cpu_initialization=config.use_cpu_initialization)
        else:
            self.embed_tokens = None
        # layers keep their GLOBAL index as module name so pipeline stages save global-numbered keys
        self.layers = nn.ModuleDict({str(i): LlamaDecoderLayer(config, i, device) for i in range(lo, hi)})
        self.norm = LlamaRMSNorm(config.hidden_size, config.rms_norm_eps, config.sequence_parallel_enabled,
                                 config.dtype, device) if with_final_norm else None
        self.rotary_emb = LlamaRotaryEmbedding(config, device)

    def position_ids(self, seq_len: int, device) -> torch.Tensor:
        cp, cp_rank = ps.get_context_model_parallel_size(), ps.get_context_model_parallel_rank()
        start = cp_rank * seq_len if cp > 1 else 0
        return torch.arange(start, start + seq_len, device=device)

    def embed(self, input_ids: torch.Tensor) -> torch.Tensor:
        h = self.embed_tokens(input_ids)
        if not self.embed_tokens.sequence_parallel_enabled:
            h = h.transpose(0, 1).contiguous()      # [B,S,H] -> [S,B,H]
        return h

    def forward(self, input_ids: Optional[torch.Tensor] = None, hidden: Optional[torch.Tensor] = None,
                seq_len: Optional[int] = None) -> torch.Tensor:
        if hidden is None:
            seq_len = input_ids.shape[1]
            hidden = self.embed(input_ids)
        assert seq_len is not None
        cos, sin = self.rotary_emb(self.position_ids(seq_len, hidden.device))
        for layer in self.layers.values():
            hidden = layer(hidden, cos, sin)
        if self.norm is not None:
            hidden = self.norm(hidden)
        return hidden


class LlamaForCausalLM(nn.Module):
    """forward(input_ids, labels) -> (loss, logits|None).  ``layer_range`` / ``with_*`` carve out a
    pipeline stage; the full model is the default."""

    def __init__(self, config: LlamaConfig, device=None, layer_range: Optional[Tuple[int, int]] = None,
                 with_embedding: bool = True, with_head: bool = True):
        super().__init__()
        self.config = config
        self.model = LlamaModel(config, device, layer_range, with_embedding, with_final_norm=with_head)
        std = config.initializer_range
        init = lambda t: nn.init.normal_(t, 0.0, std)
        if with_head:
            self.lm_head = ColumnParallelLinear(config.hidden_size, config.vocab_size, bias=False, gather_output=False,
                                                init_method=init, dtype=config.dtype, device=device,
                                                sequence_parallel_enabled=config.sequence_parallel_enabled,
                                                use_cpu_initialization=config.use_cpu_initialization)
            if config.tie_word_embeddings and with_embedding:
                self.lm_head.weight = self.model.embed_tokens.weight
                self.lm_head.weight._multi_use = True      # two gradient contributions per backward (embedding + head)
        else:
            self.lm_head = None

    # ---- loss head -------------------------------------------------------------------------------
    def _shift_labels(self, labels: torch.Tensor) -> torch.Tensor:
        if ps.get_context_model_parallel_size() > 1:
            return labels  # CP splitter hands out next-token-aligned labels (reference :817-823)
        pad = torch.full((labels.shape[0], 1), -100, dtype=labels.dtype, device=labels.device)
        return torch.cat([labels[:, 1:], pad], dim=1)

    def compute_loss(self, hidden: torch.Tensor, labels: torch.Tensor, return_logits: bool = False,
                     loss_mask: Optional[torch.Tensor] = None):
        """hidden [S_local,B,H] (post final norm) ; labels [B,S] un-shifted."""
        cfg = self.config
        tgt = self._shift_labels(labels)
        if loss_mask is not None:
            lm = self._shift_labels(loss_mask.to(labels.dtype)) if ps.get_context_model_parallel_size() == 1 else loss_mask
            tgt = tgt.masked_fill(lm <= 0, -100)
        tgt_sb = tgt.transpose(0, 1).contiguous()                      # [S,B]
        if cfg.fused_linear_ce and not return_logits:
            h = hidden
            if self.lm_head.sequence_parallel_enabled:
                h = mp.gather_from_sequence_parallel_region(h, True, self.lm_head.group)
            else:
                h = mp.copy_to_tensor_model_parallel_region(h, self.lm_head.group)
            loss_sum, n = fused_linear_cross_entropy(h.reshape(-1, h.shape[-1]), self.lm_head.weight,
                                                     tgt_sb.reshape(-1), self.lm_head.group,
                                                     chunk_tokens=cfg.ce_chunk_tokens)
            return loss_sum / n.clamp(min=1).to(loss_sum.dtype), None
        logits = self.lm_head(hidden)                                  # [S,B,V/TP]
        per_tok = parallel_cross_entropy(logits, tgt_sb, self.lm_head.group)
        n = (tgt_sb != -100).sum().clamp(min=1)
        loss = per_tok.sum() / n
        return loss, (logits.transpose(0, 1) if return_logits else None)

    def logits(self, hidden: torch.Tensor) -> torch.Tensor:
        """[B,S,V/TP] vocab-parallel logits (DPO / evaluation paths)."""
        return self.lm_head(hidden).transpose(0, 1)

    def forward(self, input_ids: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
                attention_mask: Optional[torch.Tensor] = None, hidden: Optional[torch.Tensor] = None,
                seq_len: Optional[int] = None, return_logits: bool = False, loss_mask: Optional[torch.Tensor] = None):
        hidden = self.model(input_ids=input_ids, hidden=hidden, seq_len=seq_len)
        if self.lm_head is None:
            return hidden
        if labels is None:
            return None, self.logits(hidden)
        return self.compute_loss(hidden, labels, return_logits, loss_mask)


def llama_param_count(cfg: LlamaConfig) -> int:
    h, i, v, L = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.num_hidden_layers
    kv = cfg.num_key_value_heads * cfg.head_dim
    q = cfg.num_attention_heads * cfg.head_dim
    per_layer = h * (q + 2 * kv) + q * h + 3 * h * i + 2 * h
    return L * per_layer + v * h * (1 if cfg.tie_word_embeddings else 2) + h
