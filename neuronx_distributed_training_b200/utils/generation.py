"""Greedy / sampled generation on the training model classes (used by the SFT evaluation tool).
The training models have no KV cache; evaluation prompts are short, so every step re-runs the prefix."""
from __future__ import annotations

from typing import Optional

import torch

from ..parallel import mappings as mp
from ..parallel import state as ps


@torch.no_grad()
def generate(model, input_ids: torch.Tensor, max_new_tokens: int = 64, eos_token_id: Optional[int] = None,
             temperature: float = 0.0, pad_to_multiple: int = 1) -> torch.Tensor:
    """input_ids [B, S0] -> [B, S0 + n].  Works with TP (logits are vocab-parallel and gathered)."""
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
        nxt = torch.argmax(last, -1) if temperature <= 0 else torch.multinomial(torch.softmax(last / temperature, -1), 1).squeeze(-1)
        if eos_token_id is not None:
            nxt = torch.where(finished, torch.full_like(nxt, eos_token_id), nxt)
            finished |= nxt == eos_token_id
        ids = torch.cat([ids, nxt.unsqueeze(1)], 1)
        if eos_token_id is not None and bool(finished.all()):
            break
    return ids
