"""NVLS (NVSwitch multicast) workspace: collectives issued as ``multimem.*`` instructions from our own kernels.

``ops/csrc/nvls.cu`` holds the kernels; this module owns the symmetric allocation, its multicast mapping
(``torch.distributed._symmetric_memory``: ``handle.multicast_ptr``) and the arrival-counter bookkeeping:

* :meth:`NvlsWorkspace.rmsnorm_all_gather` -- the sequence-parallel RMSNorm *is* the all-gather: every rank stores its
  normalised rows once through the multicast address and the switch replicates them into the gathered buffer of all ranks
  (reference: the norm at ``modeling_llama.py:145-161`` followed by the sequence-parallel gather of the column linear).
* :meth:`NvlsWorkspace.all_reduce` -- one-launch in-switch sum for small fp32 tensors (every rank reads the whole reduced
  vector with ``multimem.ld_reduce``), for the latency-bound tensor-parallel reductions.

Barrier protocol: two u32 arrival counters used alternately (event ``i`` uses counter ``i & 1``); every CTA of every rank
adds 1 to ALL ranks' copies with one ``multimem.red.release.sys`` and one thread per rank spins on its local copy until the
expected total is reached.  Alternating the counters makes a fast rank's arrivals for event ``i + 1`` unable to complete a
slow rank's wait for event ``i``; a rank cannot reach event ``i + 2`` before every rank has passed event ``i``.
All calls are collective over the group and must be issued in the same order on every rank (as any collective).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .. import ops

_FLAG_BYTES = 4096
_REDUCE_BYTES = 1 << 20


def multicast_available(group=None) -> bool:
    """True when a symmetric allocation of this group gets an NVSwitch multicast mapping on this machine."""
    try:
        ws = NvlsWorkspace(group, gather_bytes=1 << 16)
    except Exception:  # noqa: BLE001 - no CUDA / no symmetric memory / no multicast support
        return False
    return ws.mc_ptr != 0


class NvlsWorkspace:
    def __init__(self, group=None, gather_bytes: int = 0):
        assert torch.cuda.is_available() and dist.is_initialized()
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.gather_bytes = (int(gather_bytes) + 4095) // 4096 * 4096
        self.reduce_off = _FLAG_BYTES
        self.gather_off = [_FLAG_BYTES + _REDUCE_BYTES, _FLAG_BYTES + _REDUCE_BYTES + self.gather_bytes]
        total = self.gather_off[1] + self.gather_bytes
        try:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                symm_mem.enable_symm_mem_for_group(self.group.group_name)
        except Exception:  # noqa: BLE001 - implicit on newer torch
            pass
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=self.device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group.group_name)
        self.mc_ptr = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        self.local_ptr = int(self.buf.data_ptr())
        self.event = 0                      # barrier events issued so far
        self.count = [0, 0]                 # expected value of the two arrival counters
        self.parity = 0                     # gather-buffer double buffering
        self.reduce_parity = 0              # all-reduce staging double buffering
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        if self.mc_ptr == 0:
            raise RuntimeError("NVLS: the symmetric allocation has no multicast mapping on this machine")

    # ---- arrival counters ------------------------------------------------------------------------------------------
    def _next_event(self, arrivals_per_rank: int) -> Tuple[int, int, int]:
        """-> (multicast address, local address, value BEFORE the event) of the counter this event uses."""
        c = self.event & 1
        before = self.count[c]
        self.count[c] = (before + self.world * arrivals_per_rank) & 0xFFFFFFFF
        self.event += 1
        return self.mc_ptr + 4 * c, self.local_ptr + 4 * c, before

    # ---- norm + all-gather -----------------------------------------------------------------------------------------
    def rmsnorm_all_gather(self, x2d: torch.Tensor, weight: torch.Tensor, eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
        """x2d [m, H] bf16 (this rank's sequence shard) -> (gathered normalised [world * m, H] view of the symmetric
        buffer, rstd [m] fp32).  The view stays valid until the call after the next one (two buffers)."""
        m, H = x2d.shape
        nbytes = self.world * m * H * 2
        if nbytes > self.gather_bytes:
            raise ValueError(f"gathered tensor of {nbytes} bytes exceeds the NVLS workspace ({self.gather_bytes})")
        C = ops.C()
        off = self.gather_off[self.parity]
        self.parity ^= 1
        mc_flag, local_flag, before = self._next_event(int(C.nvls_grid_for_rows(m)))
        ops._count()
        rstd = C.rmsnorm_allgather_mc(x2d, weight, float(eps), self.mc_ptr + off, self.rank, self.world, mc_flag, local_flag,
                                      before)
        y = self.buf[off: off + nbytes].view(torch.bfloat16).view(self.world * m, H)
        return y, rstd

    # ---- small all-reduce ------------------------------------------------------------------------------------------
    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        """Sum over the group of a small fp32 tensor (<= 512 KiB), reduced inside the switch; returns a new tensor."""
        assert t.dtype == torch.float32 and t.is_cuda
        n = t.numel()
        n_pad = (n + 3) // 4 * 4
        half = _REDUCE_BYTES // 2
        if n_pad * 4 > half:
            raise ValueError("NVLS all_reduce: tensor larger than the staging region")
        src = t.reshape(-1).contiguous()
        if n_pad != n:
            src = torch.cat([src, src.new_zeros(n_pad - n)])
        out = torch.empty_like(src)
        C = ops.C()
        off = self.reduce_off + half * (self.reduce_parity & 1)      # two staging regions used alternately
        self.reduce_parity ^= 1
        mc_flag, local_flag, before = self._next_event(int(C.nvls_allreduce_arrivals()))
        ops._count()
        C.nvls_allreduce(src, out, self.local_ptr + off, self.mc_ptr + off, self.world, mc_flag, local_flag, before)
        return out[:n].view_as(t)


class _RMSNormAllGatherFn(torch.autograd.Function):
    """y_gathered = all_gather(rmsnorm(x_local)) in one NVLS kernel.  Backward: the gradient of the gathered tensor is
    reduce-scattered back to this rank's rows (every rank's consumer contributes to every row), then the ordinary
    RMSNorm backward runs on the local rows."""

    @staticmethod
    def forward(ctx, x, weight, eps, ws: NvlsWorkspace, keep: bool):
        shape = x.shape                                            # [S/tp, B, H]
        x2d = x.reshape(-1, shape[-1])
        y, rstd = ws.rmsnorm_all_gather(x2d, weight, eps)
        ctx.save_for_backward(x2d, weight, rstd)
        ctx.ws, ctx.shape = ws, shape
        out = y.view(ws.world * shape[0], *shape[1:])
        return out.clone() if keep else out

    @staticmethod
    def backward(ctx, gy):
        x2d, weight, rstd = ctx.saved_tensors
        ws = ctx.ws
        g2d = gy.reshape(-1, x2d.shape[-1]).contiguous()
        g_local = torch.empty_like(x2d)
        dist.reduce_scatter_tensor(g_local, g2d, group=ws.group)
        dx, dw = ops.C().rmsnorm_bwd(g_local, x2d, weight, rstd)
        ops._count()
        return dx.view(ctx.shape), dw.to(weight.dtype), None, None, None


def rmsnorm_all_gather(x: torch.Tensor, weight: torch.Tensor, eps: float, ws: NvlsWorkspace, keep: bool = True) -> torch.Tensor:
    """Sequence-parallel RMSNorm + all-gather along dim 0: x [S/tp, B, H] -> [S, B, H].  ``keep`` returns a private copy
    (needed when the result is saved for backward: the workspace buffer is recycled two calls later)."""
    return _RMSNormAllGatherFn.apply(x, weight, eps, ws, keep)


_WORKSPACES = {}


def get_workspace(group, gather_bytes: int) -> Optional[NvlsWorkspace]:
    """Per-group cached workspace, grown on demand; None when multicast is unavailable."""
    key = getattr(group, "group_name", "world")
    ws = _WORKSPACES.get(key)
    if ws is False:
        return None
    if ws is None or ws.gather_bytes < gather_bytes:
        try:
            ws = NvlsWorkspace(group, gather_bytes)
        except Exception:  # noqa: BLE001
            _WORKSPACES[key] = False
            return None
        _WORKSPACES[key] = ws
    return ws


def enable_nvls_norm_allgather(model: torch.nn.Module, seq_len: int, micro_batch: int, group=None) -> int:
    """Attach one NVLS workspace (tensor-parallel group) to every sequence-parallel column linear that directly follows an
    RMSNorm in a decoder layer, and switch those layers to the norm-inside-the-linear call path.  Returns the number of
    linears switched (0 when the machine has no multicast mapping: the model is left untouched)."""
    from . import state as ps
    from .layers import ColumnParallelLinear, GQAQKVColumnParallelLinear
    group = group if group is not None else ps.get_tensor_model_parallel_group()
    if dist.get_world_size(group) < 2:
        return 0
    hidden = 0
    targets = []
    for layer in model.modules():
        if not hasattr(layer, "nvls_pre_norm"):
            continue
        for holder, name in ((getattr(layer, "self_attn", None), "qkv_proj"), (getattr(layer, "mlp", None), "gate_up_proj")):
            lin = getattr(holder, name, None) if holder is not None else None
            if isinstance(lin, (ColumnParallelLinear, GQAQKVColumnParallelLinear)) and lin.sequence_parallel_enabled:
                targets.append((layer, lin))
                hidden = max(hidden, int(lin.input_size))
    if not targets:
        return 0
    ws = get_workspace(group, seq_len * micro_batch * hidden * 2)
    if ws is None:
        return 0
    for layer, lin in targets:
        lin.nvls = ws
        layer.nvls_pre_norm = True
    return len(targets)
