"""NeMo-Megatron (NNM) model checkpoint -> this framework's sharded checkpoint layout.

An NNM run stores one ``tp_rank_XX_pp_rank_YYY/model_optim_rng.ckpt`` per (TP, PP) rank whose ``state_dict`` uses
``model.language_model...`` prefixes and *stage-local* layer indices.  The framework (like the reference's NxDT layout)
expects ``model/dp_rank_00_tp_rank_XX_pp_rank_YY.pt`` shards with GLOBAL layer indices and no ``model.`` prefix, written
as a tensor-reference table plus one file per tensor so that no process ever holds more than one shard in memory.
Behaviour parity: reference examples/checkpoint_converter_scripts/nnm_model_ckpt_to_nxdt_model_ckpt_converter.py:77-152
(per-worker conversion), :19-75 (arguments, versioned output directory), :176-205 (validation, markers, process pool).
"""
from __future__ import annotations

import os
import re
import time
from concurrent.futures import ProcessPoolExecutor
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ..pipeline.partition import stage_layer_ranges
from . import io as ckpt_io

_LAYER_RE = re.compile(r"(^|\.)layers\.(\d+)\.")


def local_to_global_key(key: str, layer_offset: int) -> str:
    """``model.language_model.encoder.layers.3.x`` on a stage that starts at global layer 8 ->
    ``language_model.encoder.layers.11.x`` (only the layer index of the ``layers.N`` component moves)."""
    if key.startswith("model.language_model"):
        key = key[len("model."):]
    return _LAYER_RE.sub(lambda m: f"{m.group(1)}layers.{int(m.group(2)) + layer_offset}.", key, count=1)


def local_to_global_layers(state: Dict[str, torch.Tensor], pp_rank: int, pp_size: int, n_layers: int,
                           pipeline_cuts: Optional[Sequence[str]] = None) -> Dict[str, torch.Tensor]:
    lo = stage_layer_ranges(n_layers, pp_size, cuts=pipeline_cuts)[pp_rank][0][0]
    out = {}
    for k, v in state.items():
        nk = local_to_global_key(k, lo)
        if nk in out:
            raise ValueError(f"two NNM keys map to {nk}")
        out[nk] = v
    return out


def nnm_worker_dir(tp_rank: int, pp_rank: int) -> str:
    return f"tp_rank_{tp_rank:02d}_pp_rank_{pp_rank:03d}"


def choose_output_dir(path: str) -> str:
    """``path`` if it is missing or empty, otherwise the first free ``path_vN``."""
    if not os.path.exists(path):
        os.makedirs(path)
        return path
    if not os.listdir(path):
        return path
    n = 1
    while os.path.exists(f"{path}_v{n}"):
        n += 1
    os.makedirs(f"{path}_v{n}")
    return f"{path}_v{n}"


def convert_worker(nnm_dir: str, model_dir: str, tp_rank: int, pp_rank: int, pp_size: int, n_layers: int,
                   pipeline_cuts: Optional[Sequence[str]] = None, use_xser: bool = True) -> str:
    src = os.path.join(nnm_dir, nnm_worker_dir(tp_rank, pp_rank), "model_optim_rng.ckpt")
    ckpt = torch.load(src, map_location="cpu", weights_only=False)
    state = ckpt["state_dict"] if isinstance(ckpt, dict) and "state_dict" in ckpt else ckpt
    state = local_to_global_layers(state, pp_rank, pp_size, n_layers, pipeline_cuts)
    dst = os.path.join(model_dir, ckpt_io.shard_file_name(0, tp_rank, pp_rank))
    ckpt_io._save_object(state, dst, use_xser, async_save=False)
    return dst


def convert(nnm_ckpt_path: str, out_ckpt_path: str, tp: int, pp: int, n_layers: int, num_processes: int = 0,
            pipeline_cuts: Optional[Sequence[str]] = None, use_xser: bool = True) -> str:
    expected = {nnm_worker_dir(i, j) for i in range(tp) for j in range(pp)}
    actual = {d for d in os.listdir(nnm_ckpt_path) if os.path.isdir(os.path.join(nnm_ckpt_path, d))}
    if expected != actual:
        raise ValueError(f"{nnm_ckpt_path} does not hold exactly the TP={tp} x PP={pp} worker folders: missing "
                         f"{sorted(expected - actual)}, unexpected {sorted(actual - expected)}")
    if not pipeline_cuts and n_layers % pp != 0:
        raise ValueError(f"n_layers={n_layers} is not divisible by pp={pp}")
    out = choose_output_dir(out_ckpt_path)
    model_dir = os.path.join(out, "model")
    os.makedirs(model_dir, exist_ok=True)
    jobs: List[Tuple] = [(nnm_ckpt_path, model_dir, i, j, pp, n_layers, pipeline_cuts, use_xser)
                         for i in range(tp) for j in range(pp)]
    t0 = time.time()
    if num_processes and num_processes > 1:
        with ProcessPoolExecutor(max_workers=num_processes) as ex:
            for f in [ex.submit(convert_worker, *j) for j in jobs]:
                f.result()
    else:
        for j in jobs:
            convert_worker(*j)
    tag = os.path.basename(os.path.normpath(out))
    with open(os.path.join(out, "checkpoint"), "w") as f:     # markers last: a torn conversion is never picked up
        f.write(tag + "\n")
    with open(os.path.join(out, "done"), "w") as f:
        f.write(str(time.time()) + "\n")
    print(f"converted {len(jobs)} NNM shards to {out} in {time.time() - t0:.1f} s")
    return out
