"""Pipeline stage partitioning: even split of transformer layers or explicit ``pipeline_cuts``.

Capability parity: NxD ``NxDPPModel`` partitioning (``auto_partition`` vs
``pipeline_cuts`` as in reference conf/hf_llama3_70B_trn2_config.yaml:73).  The upstream engine is
FX-tracing based; here a stage is simply the model constructed with a ``layer_range`` (models are
written to be constructible per stage), so there is no tracer and no leaf-module bookkeeping.
"""
from __future__ import annotations

import re
from typing import List, Optional, Sequence, Tuple


def even_partition(num_layers: int, num_stages: int) -> List[Tuple[int, int]]:
    """Contiguous ranges; remainders go to the earliest stages."""
    base, rem = divmod(num_layers, num_stages)
    out, lo = [], 0
    for s in range(num_stages):
        n = base + (1 if s < rem else 0)
        out.append((lo, lo + n))
        lo += n
    return out


def cuts_partition(num_layers: int, cuts: Sequence[str]) -> List[Tuple[int, int]]:
    """``cuts`` name the LAST layer of every stage but the final one, e.g. ``model.layers.15``."""
    idx = []
    for c in cuts:
        m = re.search(r"(\d+)(?!.*\d)", str(c))
        if not m:
            raise ValueError(f"cannot parse layer index from pipeline cut {c!r}")
        idx.append(int(m.group(1)))
    if sorted(idx) != idx or len(set(idx)) != len(idx) or (idx and idx[-1] >= num_layers - 1 + 1):
        raise ValueError(f"pipeline_cuts must be strictly increasing layer names below {num_layers}: {cuts}")
    out, lo = [], 0
    for i in idx:
        out.append((lo, i + 1))
        lo = i + 1
    out.append((lo, num_layers))
    return out


def stage_layer_ranges(num_layers: int, pp: int, vpp: int = 1, cuts: Optional[Sequence[str]] = None) -> List[List[Tuple[int, int]]]:
    """ranges[pp_rank][chunk] -> (lo, hi).  With virtual pipelining the model is cut into pp*vpp chunks and
    rank r owns chunks r, r+pp, r+2pp, ... (Megatron interleaving)."""
    n_chunks = pp * vpp
    if cuts:
        parts = cuts_partition(num_layers, cuts)
        if len(parts) != n_chunks:
            raise ValueError(f"{len(cuts)} pipeline_cuts give {len(parts)} stages, expected {n_chunks}")
    else:
        parts = even_partition(num_layers, n_chunks)
    return [[parts[c * pp + r] for c in range(vpp)] for r in range(pp)]
