"""Process-group grid: world = PP x DP x CP x TP with TP fastest varying (adjacent ranks share a
TP group, so on an 8-GPU NVSwitch node TP traffic never leaves the node).

Capability parity with NxD ``parallel_state`` as used by the reference (call
sites e.g. reference lightning_modules/model/base.py:232,388-394,427-450 and the layout
illustrated in models/megatron/megatron_init.py:104-119).  Rank decomposition:

    rank = ((pp_rank * DP + dp_rank) * CP + cp_rank) * TP + tp_rank

Expert parallelism carves EP-sized groups out of the data-parallel dimension
(expert-model-parallel group) plus the complementary expert-data-parallel group.
Works without ``torch.distributed`` being initialised (single process => every group has size 1).
"""
from __future__ import annotations

import datetime
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


@dataclass
class _Grid:
    world_size: int = 1
    rank: int = 0
    tp: int = 1
    pp: int = 1
    cp: int = 1
    dp: int = 1
    ep: int = 1
    vpp: int = 1
    tp_rank: int = 0
    pp_rank: int = 0
    cp_rank: int = 0
    dp_rank: int = 0
    ep_rank: int = 0
    vpp_rank: int = 0
    initialized: bool = False


_G = _Grid()
_GROUPS: Dict[str, Optional[dist.ProcessGroup]] = {}
_GROUP_RANKS: Dict[str, List[int]] = {}


def init_distributed(backend: Optional[str] = None, timeout_s: int = 1800) -> None:
    """Initialise torch.distributed from torchrun-style env vars (no-op if single process)."""
    if dist.is_initialized():
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and "RANK" not in os.environ:
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", timeout=datetime.timedelta(seconds=timeout_s),
                                device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group(backend=backend, timeout=datetime.timedelta(seconds=timeout_s))


def rank_coords(rank: int, tp: int, cp: int, dp: int, pp: int) -> Tuple[int, int, int, int]:
    """(pp_rank, dp_rank, cp_rank, tp_rank) of a global rank."""
    tp_rank = rank % tp
    cp_rank = (rank // tp) % cp
    dp_rank = (rank // (tp * cp)) % dp
    pp_rank = rank // (tp * cp * dp)
    return pp_rank, dp_rank, cp_rank, tp_rank


def coords_to_rank(pp_rank: int, dp_rank: int, cp_rank: int, tp_rank: int, tp: int, cp: int, dp: int) -> int:
    return ((pp_rank * dp + dp_rank) * cp + cp_rank) * tp + tp_rank


def enumerate_groups(world: int, tp: int, pp: int, cp: int, ep: int = 1) -> Dict[str, List[List[int]]]:
    """All rank lists of every group kind (pure function; unit-testable without dist)."""
    dp = world // (tp * pp * cp)
    out: Dict[str, List[List[int]]] = {k: [] for k in
                                       ("tp", "cp", "dp", "pp", "dp_cp", "mp", "ep", "edp", "tp_cp", "embed")}
    for p in range(pp):
        for d in range(dp):
            for c in range(cp):
                out["tp"].append([coords_to_rank(p, d, c, t, tp, cp, dp) for t in range(tp)])
            out["tp_cp"].append([coords_to_rank(p, d, c, t, tp, cp, dp) for c in range(cp) for t in range(tp)])
        for d in range(dp):
            for t in range(tp):
                out["cp"].append([coords_to_rank(p, d, c, t, tp, cp, dp) for c in range(cp)])
        for c in range(cp):
            for t in range(tp):
                out["dp"].append([coords_to_rank(p, d, c, t, tp, cp, dp) for d in range(dp)])
                # expert groups split the DP dimension: consecutive EP ranks form an expert-parallel
                # group; ranks with equal (d % ep) form the expert-data-parallel group.
                for base in range(0, dp, ep):
                    out["ep"].append([coords_to_rank(p, d, c, t, tp, cp, dp) for d in range(base, base + ep)])
                for off in range(ep):
                    out["edp"].append([coords_to_rank(p, d, c, t, tp, cp, dp) for d in range(off, dp, ep)])
        for t in range(tp):
            out["dp_cp"].append([coords_to_rank(p, d, c, t, tp, cp, dp) for d in range(dp) for c in range(cp)])
    for d in range(dp):
        for c in range(cp):
            for t in range(tp):
                ranks = [coords_to_rank(p, d, c, t, tp, cp, dp) for p in range(pp)]
                out["pp"].append(ranks)
                out["embed"].append([ranks[0], ranks[-1]] if pp > 1 else [ranks[0]])
            # model-parallel group = TP x PP for a fixed (dp, cp)
            out["mp"].append([coords_to_rank(p, d, c, t, tp, cp, dp) for p in range(pp) for t in range(tp)])
    return out


def initialize_model_parallel(tensor_model_parallel_size: int = 1, pipeline_model_parallel_size: int = 1,
                              context_parallel_size: int = 1, expert_model_parallel_size: int = 1,
                              virtual_pipeline_model_parallel_size: int = 1) -> None:
    global _G
    destroy_model_parallel()
    if dist.is_initialized():
        world, rank = dist.get_world_size(), dist.get_rank()
    else:
        world, rank = 1, 0
    tp, pp, cp, ep = (int(tensor_model_parallel_size), int(pipeline_model_parallel_size),
                      int(context_parallel_size), int(expert_model_parallel_size))
    if world % (tp * pp * cp) != 0:
        raise ValueError(f"world size {world} not divisible by tp*pp*cp = {tp}*{pp}*{cp}")
    dp = world // (tp * pp * cp)
    if dp % ep != 0:
        raise ValueError(f"expert_model_parallel_size {ep} must divide data parallel size {dp}")
    pp_rank, dp_rank, cp_rank, tp_rank = rank_coords(rank, tp, cp, dp, pp)
    _G = _Grid(world_size=world, rank=rank, tp=tp, pp=pp, cp=cp, dp=dp, ep=ep,
               vpp=int(virtual_pipeline_model_parallel_size or 1),
               tp_rank=tp_rank, pp_rank=pp_rank, cp_rank=cp_rank, dp_rank=dp_rank,
               ep_rank=dp_rank % ep, initialized=True)
    groups = enumerate_groups(world, tp, pp, cp, ep)
    for kind, rank_lists in groups.items():
        mine = None
        for ranks in rank_lists:
            # every process must call new_group for every group, in the same order
            g = None
            if dist.is_initialized() and world > 1:
                g = dist.new_group(ranks=ranks) if len(ranks) > 1 else None
            if rank in ranks:
                mine = (g, ranks)
        if mine is None:
            # only first/last pipeline stages belong to an embedding group
            assert kind == "embed", f"rank {rank} is in no {kind} group"
            mine = (None, [])
        _GROUPS[kind], _GROUP_RANKS[kind] = mine


_SHUFFLE_GROUPS: Dict[int, Tuple[object, List[int]]] = {}


def get_token_shuffle_group(size: int):
    """Process group of ``size`` consecutive data-parallel ranks (MoE ``token_shuffle_group_size``), created on first
    use.  Every rank of the world must call this at the same point (groups are created collectively)."""
    size = int(size)
    if size <= 1 or not dist.is_initialized() or _G.world_size == 1:
        return None, []
    if size in _SHUFFLE_GROUPS:
        return _SHUFFLE_GROUPS[size]
    if _G.dp % size != 0:
        raise ValueError(f"token_shuffle_group_size {size} must divide the data-parallel size {_G.dp}")
    mine = (None, [])
    for ranks in enumerate_groups(_G.world_size, _G.tp, _G.pp, _G.cp, _G.ep)["dp"]:
        for i in range(0, len(ranks), size):
            sub = ranks[i:i + size]
            g = dist.new_group(ranks=sub)
            if _G.rank in sub:
                mine = (g, sub)
    _SHUFFLE_GROUPS[size] = mine
    return mine


_KV_REPLICA_GROUPS: Dict[int, Tuple[object, List[int]]] = {}


def get_kv_replica_group(reps: int):
    """Process group of ``reps`` consecutive tensor-parallel ranks that hold replicas of the same logical KV head
    (``kv_size_multiplier`` / ``kv_replicator``).  ONE set of groups per (layout, reps) for the whole model: every
    attention layer looks the shared group up here, so the number of communicators does not grow with the layer
    count and pipeline stages with different layer counts issue identical ``new_group`` sequences (the groups are
    created collectively on first use; every rank builds at least one attention layer)."""
    reps = int(reps)
    if reps <= 1 or not dist.is_initialized() or _G.world_size == 1:
        return None, []
    if reps in _KV_REPLICA_GROUPS:
        return _KV_REPLICA_GROUPS[reps]
    if _G.tp % reps != 0:
        raise ValueError(f"KV replica group size {reps} must divide the tensor-parallel size {_G.tp}")
    mine = (None, [])
    for ranks in enumerate_groups(_G.world_size, _G.tp, _G.pp, _G.cp, _G.ep)["tp"]:
        for i in range(0, len(ranks), reps):
            sub = ranks[i:i + reps]
            g = dist.new_group(ranks=sub)
            if _G.rank in sub:
                mine = (g, sub)
    _KV_REPLICA_GROUPS[reps] = mine
    return mine


def destroy_model_parallel() -> None:
    _SHUFFLE_GROUPS.clear()
    _KV_REPLICA_GROUPS.clear()
    global _G
    _GROUPS.clear()
    _GROUP_RANKS.clear()
    _G = _Grid()


def model_parallel_is_initialized() -> bool:
    return _G.initialized


def _group(kind: str):
    return _GROUPS.get(kind)


# ---- getters (names follow the capability surface the reference uses) -------------------------------
def get_world_size() -> int: return _G.world_size
def get_rank() -> int: return _G.rank

def get_tensor_model_parallel_size() -> int: return _G.tp
def get_tensor_model_parallel_rank() -> int: return _G.tp_rank
def get_tensor_model_parallel_group(): return _group("tp")
def get_tensor_model_parallel_ranks() -> List[int]: return _GROUP_RANKS.get("tp", [0])
def get_tensor_model_parallel_src_rank() -> int: return _GROUP_RANKS.get("tp", [0])[0]

def get_pipeline_model_parallel_size() -> int: return _G.pp
def get_pipeline_model_parallel_rank() -> int: return _G.pp_rank
def get_pipeline_model_parallel_group(): return _group("pp")
def get_pipeline_model_parallel_ranks() -> List[int]: return _GROUP_RANKS.get("pp", [0])
def is_pipeline_first_stage() -> bool: return _G.pp_rank == 0
def is_pipeline_last_stage() -> bool: return _G.pp_rank == _G.pp - 1
def get_pipeline_model_parallel_next_rank() -> int:
    r = _GROUP_RANKS.get("pp", [0]); return r[(_G.pp_rank + 1) % len(r)]
def get_pipeline_model_parallel_prev_rank() -> int:
    r = _GROUP_RANKS.get("pp", [0]); return r[(_G.pp_rank - 1) % len(r)]
def get_virtual_pipeline_model_parallel_size() -> int: return _G.vpp
def get_virtual_pipeline_model_parallel_rank() -> int: return _G.vpp_rank
def set_virtual_pipeline_model_parallel_rank(r: int) -> None: _G.vpp_rank = r

def get_data_parallel_size() -> int: return _G.dp
def get_data_parallel_rank() -> int: return _G.dp_rank
def get_data_parallel_group(): return _group("dp")
def get_data_parallel_ranks() -> List[int]: return _GROUP_RANKS.get("dp", [0])
def get_data_context_parallel_group(): return _group("dp_cp")
def get_data_context_parallel_ranks() -> List[int]: return _GROUP_RANKS.get("dp_cp", [0])

def get_context_model_parallel_size() -> int: return _G.cp
def get_context_model_parallel_rank() -> int: return _G.cp_rank
def get_context_model_parallel_group(): return _group("cp")
def get_context_model_parallel_ranks() -> List[int]: return _GROUP_RANKS.get("cp", [0])

def get_expert_model_parallel_size() -> int: return _G.ep
def get_expert_model_parallel_rank() -> int: return _G.ep_rank
def get_expert_model_parallel_group(): return _group("ep")
def get_expert_data_parallel_group(): return _group("edp")
def get_expert_data_parallel_size() -> int: return _G.dp // _G.ep

def get_model_parallel_group(): return _group("mp")
def get_embedding_group(): return _group("embed")
def get_embedding_ranks() -> List[int]: return _GROUP_RANKS.get("embed", [0])


def get_context_model_parallel_src_tgt_pairs() -> List[Tuple[int, int]]:
    """Ring (src, tgt) pairs over this rank's CP group: KV blocks travel rank i -> i+1."""
    ranks = get_context_model_parallel_ranks()
    n = len(ranks)
    return [(ranks[i], ranks[(i + 1) % n]) for i in range(n)]


def group_size(group) -> int:
    if group is None or not dist.is_initialized():
        return 1
    return dist.get_world_size(group)


def group_rank(group) -> int:
    if group is None or not dist.is_initialized():
        return 0
    return dist.get_rank(group)


def all_reduce_scalar(value: torch.Tensor, group, op=None) -> torch.Tensor:
    if group is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(value, op=op or dist.ReduceOp.SUM, group=group)
    return value


def barrier(group=None) -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier(group=group)
