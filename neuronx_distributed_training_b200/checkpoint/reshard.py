"""Re-shard ZeRO-1 optimizer state when the data-parallel size changes between save and resume.

The flat optimizer (``optim/distributed.py``) lays every parameter out at the same offset for any DP size; only the
padded total and the shard boundaries depend on DP.  A checkpoint written by ``dp_old`` ranks therefore holds
``dp_old`` contiguous slices of one flat state per (tp, pp) coordinate, and the state of any new slice
``[lo, hi)`` is assembled by copying the overlapping pieces.  TP / PP changes need the full-model converter instead
(``checkpoint/converter.py``): the parameter shards themselves differ.

The reference stores per-DP-rank optimizer shards in the same way (docs/developer_guides/migration_nnm_nxdt.rst:138-145)
but cannot change DP on resume; this is an extension.
"""
from __future__ import annotations

import glob
import os
from typing import Any, Dict, List, Optional

import torch

from . import io as ckpt_io


def _shard_files(checkpoint_dir: str, tag: str, tp_rank: int, pp_rank: int) -> List[str]:
    pat = os.path.join(checkpoint_dir, tag, "optim", f"dp_rank_*_tp_rank_{tp_rank:02d}_pp_rank_{pp_rank:02d}.pt")
    return sorted(glob.glob(pat))


def reshard_optimizer_state(checkpoint_dir: str, tag: str, tp_rank: int, pp_rank: int, new_lo: int, new_hi: int,
                            new_total: int, expected_layout: Optional[list] = None) -> Dict[str, Any]:
    """State dict for the flat-buffer range ``[new_lo, new_hi)`` of a ``new_total``-element layout, assembled from
    every DP shard stored under ``<checkpoint_dir>/<tag>/optim`` for this (tp, pp) coordinate."""
    files = _shard_files(checkpoint_dir, tag, tp_rank, pp_rank)
    if not files:
        raise FileNotFoundError(f"no optimizer shards for tp={tp_rank} pp={pp_rank} under {checkpoint_dir}/{tag}/optim")
    n = new_hi - new_lo
    out: Dict[str, Any] = {}
    bufs: Dict[str, Optional[torch.Tensor]] = {"exp_avg": None, "exp_avg_sq": None, "master": None}
    covered = 0
    seen_dp = None
    for f in files:
        sd = ckpt_io._load_object(f, "cpu")
        if expected_layout is not None and [tuple(x) for x in sd["layout"]] != [tuple(x) for x in expected_layout]:
            raise ValueError("optimizer checkpoint was written for a different parameter layout (TP / PP / model "
                             "changed?); convert the checkpoint with checkpoint_converter.py instead of resharding")
        lo, hi = int(sd["shard_lo"]), int(sd["shard_hi"])
        seen_dp = int(sd.get("dp_size", 1))
        if not out:
            out = {k: v for k, v in sd.items() if k not in ("exp_avg", "exp_avg_sq", "master", "shard_lo", "shard_hi",
                                                            "total", "dp_size", "dp_rank")}
        a, b = max(lo, new_lo), min(hi, new_hi)
        if a >= b:
            continue
        for key in bufs:
            src = sd.get(key)
            if src is None:
                continue
            if bufs[key] is None:
                bufs[key] = torch.zeros(n, dtype=src.dtype)
            bufs[key][a - new_lo:b - new_lo] = src[a - lo:b - lo]
        covered += b - a
    # elements past the old padded total are alignment padding (never touched by the update): zeros are exact
    old_total = max(int(ckpt_io._load_object(f, "cpu")["total"]) for f in files[:1])
    need = max(0, min(new_hi, old_total) - new_lo)
    if covered < need:
        raise ValueError(f"optimizer shards under {checkpoint_dir}/{tag} cover {covered} of the {need} elements needed "
                         f"(found {len(files)} shard files, checkpoint dp_size={seen_dp}); is the checkpoint complete?")
    out.update(bufs)
    out.update({"shard_lo": new_lo, "shard_hi": new_hi, "total": new_total})
    return out


def load_optimizer_state_any_dp(optimizer, state: Optional[Dict[str, Any]], checkpoint_dir: str, tag: str,
                                tp_rank: int, pp_rank: int) -> bool:
    """``optimizer.load_state_dict`` that transparently reshards when the stored DP layout differs.
    Returns True when resharding was needed."""
    def fits(sd):
        return sd is not None and int(sd["total"]) == optimizer.total and int(sd["shard_lo"]) == optimizer.shard_lo \
            and int(sd["shard_hi"]) == optimizer.shard_hi
    if fits(state):
        optimizer.load_state_dict(state)
        return False
    layout = [(o, int(p.numel())) for o, p in zip(optimizer.offsets, optimizer.params)]
    sd = reshard_optimizer_state(checkpoint_dir, tag, tp_rank, pp_rank, optimizer.shard_lo, optimizer.shard_hi,
                                 optimizer.total, expected_layout=layout)
    optimizer.load_state_dict(sd)
    return True
