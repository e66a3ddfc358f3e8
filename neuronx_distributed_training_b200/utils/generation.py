"""Batched, KV-cached generation on the training model classes (used by the SFT evaluation tool).

The reference evaluates fine-tuned weights with separate inference runtimes (examples/sft_evaluation/evaluate.py:55-147,
models/nxd_llama.py); here the TP-sharded training modules themselves are decoded: a prefill pass fills a per-layer
key/value cache, then every new token costs one single-position forward.  Prompts of different lengths are batched with
LEFT padding; RoPE attention only depends on relative positions, so shifting a row's positions by its pad count is exact,
and padded keys are masked out.  Works for the HF-Llama and HF-Mixtral families (any decoder layer exposing
``self_attn`` (LlamaAttention), ``input_layernorm``, ``post_attention_layernorm`` and ``mlp``), with tensor parallelism.
"""
from __future__ import annotations

import contextlib
from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F

from .. import ops
from ..parallel import mappings as mp
from ..parallel import state as ps


@torch.no_grad()
def filter_logits(logits: torch.Tensor, top_k: int = 0, top_p: float = 1.0) -> torch.Tensor:
    """[B, V] logits with everything outside the top-k / nucleus (top-p) set pushed to -inf."""
    if top_k and top_k > 0:
        kth = torch.topk(logits, min(int(top_k), logits.shape[-1]), dim=-1).values[..., -1:]
        logits = logits.masked_fill(logits < kth, float("-inf"))
    if top_p < 1.0:
        srt, idx = torch.sort(logits, dim=-1, descending=True)
        cum = torch.softmax(srt, -1).cumsum(-1)
        drop = cum - torch.softmax(srt, -1) >= top_p            # tokens that start beyond the nucleus (the first always stays)
        srt = srt.masked_fill(drop, float("-inf"))
        logits = torch.full_like(logits, float("-inf")).scatter(-1, idx, srt)
    return logits


def _sample(last: torch.Tensor, temperature: float, top_k: int, top_p: float) -> torch.Tensor:
    if temperature <= 0:
        return torch.argmax(last, -1)
    return torch.multinomial(torch.softmax(filter_logits(last / temperature, top_k, top_p), -1), 1).squeeze(-1)


@contextlib.contextmanager
def sequence_parallel_disabled(model):
    """Run the tensor-parallel layers in their all-reduce form for the duration of the block: decoding feeds one
    position at a time, which cannot be split over the sequence-parallel ranks.  Mathematically the same model."""
    flipped = []
    for m in model.modules():
        for attr in ("sequence_parallel_enabled", "sp"):
            if isinstance(getattr(m, attr, None), bool) and getattr(m, attr):
                setattr(m, attr, False)
                flipped.append((m, attr))
    try:
        yield
    finally:
        for m, attr in flipped:
            setattr(m, attr, True)


def supports_kv_cache(model) -> bool:
    inner = getattr(model, "model", None)
    layers = getattr(inner, "layers", None)
    if inner is None or layers is None or getattr(model, "lm_head", None) is None or getattr(inner, "embed_tokens", None) is None:
        return False
    first = next(iter(layers.values())) if hasattr(layers, "values") else layers[0]
    return all(hasattr(first, a) for a in ("self_attn", "input_layernorm", "post_attention_layernorm", "mlp")) and \
        hasattr(first.self_attn, "qkv_mode")


class _LayerCache:
    __slots__ = ("k", "v", "len")

    def __init__(self, B, T, heads, dim, dtype, device):
        self.k = torch.empty(B, T, heads, dim, dtype=dtype, device=device)
        self.v = torch.empty(B, T, heads, dim, dtype=dtype, device=device)
        self.len = 0


def _attention_step(attn, hidden: torch.Tensor, cos, sin, cache: _LayerCache, key_valid: torch.Tensor) -> torch.Tensor:
    """hidden [S, B, H] for cache slots [cache.len, cache.len + S); key_valid [B, T_max] marks non-pad slots."""
    if attn.qkv_mode == "gqa":
        q, k, v = attn.qkv_proj(hidden)
    elif attn.qkv_mode == "fused":
        q, k, v = attn.qkv_proj(hidden).chunk(3, dim=-1)
    else:
        q, k, v = attn.q_proj(hidden), attn.k_proj(hidden), attn.v_proj(hidden)
    S, B = q.shape[0], q.shape[1]
    q = q.reshape(S, B, attn.local_q_heads, attn.head_dim).transpose(0, 1)
    k = k.reshape(S, B, attn.local_kv_heads, attn.head_dim).transpose(0, 1)
    v = v.reshape(S, B, attn.local_kv_heads, attn.head_dim).transpose(0, 1)
    q, k = ops.rope(q, k, cos, sin)
    t0 = cache.len
    cache.k[:, t0:t0 + S] = k
    cache.v[:, t0:t0 + S] = v
    cache.len = T = t0 + S
    qi = torch.arange(t0, T, device=hidden.device).view(1, S, 1)
    ki = torch.arange(T, device=hidden.device).view(1, 1, T)
    allowed = (ki <= qi) & key_valid[:, None, :T]
    window = attn.core_attn.window
    if window:
        allowed &= ki > qi - int(window)
    allowed |= ki == qi                    # a pad query attends to itself (finite garbage that no real token reads)
    rep = attn.local_q_heads // attn.local_kv_heads
    kk, vv = cache.k[:, :T].transpose(1, 2), cache.v[:, :T].transpose(1, 2)
    if rep > 1:
        kk, vv = kk.repeat_interleave(rep, dim=1), vv.repeat_interleave(rep, dim=1)
    o = F.scaled_dot_product_attention(q.transpose(1, 2), kk, vv, attn_mask=allowed[:, None], scale=attn.core_attn.scale)
    o = o.transpose(1, 2).reshape(B, S, attn.local_q_heads * attn.head_dim).transpose(0, 1)
    return attn.o_proj(o.contiguous())


def _forward_cached(model, ids: torch.Tensor, caches: List[_LayerCache], key_valid: torch.Tensor) -> torch.Tensor:
    """ids [B, S] occupying cache slots [len, len + S) -> vocab-parallel logits of the LAST position [B, V_local] (fp32)."""
    inner = model.model
    t0 = caches[0].len
    hidden = inner.embed(ids)                                               # [S, B, H]
    cos, sin = inner.rotary_emb(torch.arange(t0, t0 + ids.shape[1], device=ids.device))
    for layer, cache in zip(inner.layers.values(), caches):
        hidden = hidden + _attention_step(layer.self_attn, layer.input_layernorm(hidden), cos, sin, cache, key_valid)
        m = layer.mlp(layer.post_attention_layernorm(hidden))
        hidden = hidden + (m[0] if isinstance(m, tuple) else m)             # MoE blocks also return router logits
    hidden = inner.norm(hidden[-1:])
    return model.lm_head(hidden)[0].float()


@torch.no_grad()
def generate_cached(model, prompts: Sequence[Sequence[int]], max_new_tokens: int = 64, eos_token_id: Optional[int] = None,
                    temperature: float = 0.0, top_k: int = 0, top_p: float = 1.0, pad_token_id: int = 0) -> List[List[int]]:
    """Batched decoding of variable-length prompts; returns the NEW tokens of every prompt (eos excluded)."""
    assert supports_kv_cache(model), "generate_cached needs a Llama-family decoder (see supports_kv_cache)"
    model.eval()
    dev = next(model.parameters()).device
    dtype = next(model.parameters()).dtype
    B = len(prompts)
    L = max(len(p) for p in prompts)
    T = L + max_new_tokens
    ids = torch.full((B, L), pad_token_id, dtype=torch.long, device=dev)
    key_valid = torch.ones(B, T, dtype=torch.bool, device=dev)
    for b, p in enumerate(prompts):
        ids[b, L - len(p):] = torch.as_tensor(list(p), dtype=torch.long, device=dev)
        key_valid[b, :L - len(p)] = False
    tp = ps.get_tensor_model_parallel_size()
    attn0 = next(iter(model.model.layers.values())).self_attn
    caches = [_LayerCache(B, T, attn0.local_kv_heads, attn0.head_dim, dtype, dev) for _ in model.model.layers]
    out: List[List[int]] = [[] for _ in range(B)]
    finished = torch.zeros(B, dtype=torch.bool, device=dev)
    with sequence_parallel_disabled(model):
        cur = ids
        for step in range(max_new_tokens):
            last = _forward_cached(model, cur, caches, key_valid)
            if tp > 1:
                last = mp.all_gather_last_dim(last)
            nxt = _sample(last, temperature, top_k, top_p)
            if eos_token_id is not None:
                nxt = torch.where(finished, torch.full_like(nxt, eos_token_id), nxt)
                finished |= nxt == eos_token_id
            for b, tkn in enumerate(nxt.tolist()):
                if eos_token_id is None or tkn != eos_token_id:
                    out[b].append(tkn)
            if eos_token_id is not None and bool(finished.all()):
                break
            cur = nxt.unsqueeze(1)
    return out


@torch.no_grad()
def generate_uncached(model, input_ids: torch.Tensor, max_new_tokens: int = 64, eos_token_id: Optional[int] = None,
                      temperature: float = 0.0, pad_to_multiple: int = 1, top_k: int = 0, top_p: float = 1.0) -> torch.Tensor:
    """Reference decoder: re-runs the whole prefix for every new token through the unmodified training forward.  Used as
    the oracle for the cached path and for model classes without ``supports_kv_cache``."""
    model.eval()
    tp = ps.get_tensor_model_parallel_size()
    ids = input_ids
    finished = torch.zeros(ids.shape[0], dtype=torch.bool, device=ids.device)
    for _ in range(max_new_tokens):
        cur = ids
        pad = (-cur.shape[1]) % max(pad_to_multiple, tp if getattr(model, "config", None) is not None and
                                    getattr(model.config, "sequence_parallel_enabled", False) else 1)
        if pad:
            cur = torch.cat([cur, cur.new_zeros(cur.shape[0], pad)], 1)
        _, logits = model(input_ids=cur, labels=None)
        last = logits[:, ids.shape[1] - 1].float()
        if tp > 1:
            last = mp.all_gather_last_dim(last)
        nxt = _sample(last, temperature, top_k, top_p)
        if eos_token_id is not None:
            nxt = torch.where(finished, torch.full_like(nxt, eos_token_id), nxt)
            finished |= nxt == eos_token_id
        ids = torch.cat([ids, nxt.unsqueeze(1)], 1)
        if eos_token_id is not None and bool(finished.all()):
            break
    return ids


def generate(model, input_ids: torch.Tensor, max_new_tokens: int = 64, eos_token_id: Optional[int] = None,
             temperature: float = 0.0, pad_to_multiple: int = 1, top_k: int = 0, top_p: float = 1.0) -> torch.Tensor:
    """input_ids [B, S0] (equal-length prompts) -> [B, S0 + n].  KV-cached when the model family supports it; rows that
    finished early are filled with ``eos_token_id``."""
    if not supports_kv_cache(model):
        return generate_uncached(model, input_ids, max_new_tokens, eos_token_id, temperature, pad_to_multiple, top_k, top_p)
    new = generate_cached(model, input_ids.tolist(), max_new_tokens, eos_token_id, temperature, top_k, top_p)
    n = max((len(x) for x in new), default=0)
    if eos_token_id is not None and any(len(x) < max_new_tokens for x in new):
        n = min(max_new_tokens, n + 1)                   # room for the terminating eos
    fill = eos_token_id if eos_token_id is not None else 0
    tail = torch.full((input_ids.shape[0], n), fill, dtype=input_ids.dtype, device=input_ids.device)
    for b, x in enumerate(new):
        tail[b, :len(x)] = torch.as_tensor(x, dtype=input_ids.dtype, device=input_ids.device)
    return torch.cat([input_ids, tail], 1)
