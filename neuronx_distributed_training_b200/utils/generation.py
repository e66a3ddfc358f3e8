"""Greedy / sampled generation on the training model classes (used by the SFT evaluation tool).
The training models have no KV cache; evaluation prompts are short, so every step re-runs the prefix."""
from __future__ import annotations

from typing import Optional

import torch

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


def generate(model, input_ids: torch.Tensor, max_new_tokens: int = 64, eos_token_id: Optional[int] = None,
             temperature: float = 0.0, pad_to_multiple: int = 1, top_k: int = 0, top_p: float = 1.0) -> torch.Tensor:
    """input_ids [B, S0] -> [B, S0 + n].  Works with TP (logits are vocab-parallel and gathered).  ``temperature <= 0``:
    greedy; otherwise sampling, optionally restricted to the ``top_k`` most likely tokens and / or the smallest set whose
    probability mass reaches ``top_p``.  The forward pass is re-run over the whole prefix for every new token (no KV
    cache): adequate for the short generations of the SFT evaluation, not a serving path."""
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
        if temperature <= 0:
            nxt = torch.argmax(last, -1)
        else:
            nxt = torch.multinomial(torch.softmax(filter_logits(last / temperature, top_k, top_p), -1), 1).squeeze(-1)
        if eos_token_id is not None:
            nxt = torch.where(finished, torch.full_like(nxt, eos_token_id), nxt)
            finished |= nxt == eos_token_id
        ids = torch.cat([ids, nxt.unsqueeze(1)], 1)
        if eos_token_id is not None and bool(finished.all()):
            break
    return ids
