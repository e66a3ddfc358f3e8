"""Host side of the fused tensor-parallel kernels (``ops/csrc/fused_tp.cu``).

One ``FusedTP`` object per tensor-parallel group owns a *symmetric* workspace (same layout on every rank,
peer-mapped over NVLink through ``torch.distributed._symmetric_memory``):

    [ handshake slots | RS arrival counters x2 | AG ready flags | gathered buffers x2 | RS staging slabs x2 ]

and a monotonically increasing ``epoch`` (every fused call on every rank of the group advances it in
lock-step, which is what makes the double-buffer parity protocol safe - see the kernel file).
It offers the four conjugate ops the Column/Row parallel linears need with sequence parallelism:

    all_gather_gemm(x_local, W)              Y  = AG(X) W^T           (column fwd, row dgrad with trans_b=False)
    gemm_reduce_scatter(A, W)                Z  = RS(A W^T)           (row fwd, column dgrad with trans_b=False)
    all_gather_wgrad(dY, x_local, W)         dW += dY^T AG(X)         (column wgrad; gather fused into the K loop)
    all_gather(x_local)                      AG(X)                    (comm CTAs only)
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from .. import ops
from . import state as ps

_HS_OFF = 0
_CNT_OFF = 4096
_CNT_BYTES = 4096            # per parity: up to 1024 row-blocks of 128 rows
_FLAG_OFF = _CNT_OFF + 2 * _CNT_BYTES
_FLAG_BYTES = 16384          # ready flags of the gathered buffer (one uint32 per 128-row block), written by the peers
_DATA_OFF = _FLAG_OFF + _FLAG_BYTES


def _align(n: int, a: int = 4096) -> int:
    return (n + a - 1) // a * a


class FusedTP:
    def __init__(self, group, max_shard_bytes: int, max_stage_bytes: int, num_comm_ctas: int = 0):
        assert torch.cuda.is_available() and dist.is_initialized()
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.tp = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        assert 2 <= self.tp <= 8, "fused TP kernels support 2..8 ranks per group"
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_comm_ctas_fixed = int(num_comm_ctas or os.environ.get("NXDT_AG_COMM_CTAS", 0))
        self.num_comm_ctas = self.num_comm_ctas_fixed or (12 if self.tp >= 4 else 8)
        self.max_shard_bytes = _align(max_shard_bytes)
        self.max_gather_bytes = self.max_shard_bytes * self.tp
        self.max_stage_bytes = _align(max_stage_bytes)
        self.xg_off = [_DATA_OFF, _DATA_OFF + self.max_gather_bytes]
        x0 = _DATA_OFF + 2 * self.max_gather_bytes
        self.x_off = [x0, x0 + self.max_shard_bytes]           # pull mode: my shard staged where the peers can read it
        s0 = x0 + 2 * self.max_shard_bytes
        self.stage_off = [s0, s0 + self.max_stage_bytes]
        self.push = os.environ.get("NXDT_AG_MODE", "pull") == "push"
        total = self.stage_off[1] + self.max_stage_bytes
        try:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")     # newer torch enables every group implicitly and deprecates the call
                symm_mem.enable_symm_mem_for_group(self.group.group_name)
        except Exception:  # noqa: BLE001 - absent / no-op on newer torch
            pass
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=self.device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group.group_name)
        self.base_ptrs: List[int] = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(self.base_ptrs) == self.tp
        self.epoch = 0
        # The protocol epoch lives on the device and is advanced by a one-thread kernel in front of every fused op, so
        # a captured CUDA graph advances it on every replay (NXDT_FUSED_DEVICE_EPOCH=0: host-passed epoch, eager only).
        self.device_epoch = os.environ.get("NXDT_FUSED_DEVICE_EPOCH", "1") != "0" and hasattr(ops.C(), "tp_barrier")
        self.epoch_dev = torch.zeros(4, dtype=torch.int32, device=self.device)
        self._epoch_ptr = int(self.epoch_dev.data_ptr()) if self.device_epoch else 0
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    # ---- helpers ---------------------------------------------------------------------------------
    def _next_epoch(self) -> Tuple[int, int]:
        self.epoch += 1
        return self.epoch, self.epoch & 1

    def graph_boundary(self) -> None:
        """All-peer barrier on the current stream + restart of the staging parity.  Called at the start and at the end
        of a captured region: inside a graph the buffer parities are baked in, so consecutive replays (or a replay
        after eager calls) must not overlap across ranks."""
        if not self.device_epoch:
            raise RuntimeError("CUDA-graph capture of the fused TP kernels needs the device-side epoch")
        if self.push:
            raise RuntimeError("CUDA-graph capture of the fused TP kernels needs NXDT_AG_MODE=pull")
        hs_local, hs_peer = self._hs()
        ops._count()
        ops.C().tp_barrier(hs_local, hs_peer, self.rank, self.tp, self._epoch_ptr)
        self.epoch = 0

    def _hs(self) -> Tuple[int, List[int]]:
        return self.base_ptrs[self.rank] + _HS_OFF, [p + _HS_OFF for p in self.base_ptrs]

    def _gather_buffer(self, x2d: torch.Tensor, parity: int):
        """My symmetric gathered buffer for this call (peers push their shards into it) with my own shard copied
        in; returns (view [tp*m, K], per-rank base pointers of the same buffer, per-rank flag pointers)."""
        m, K = x2d.shape
        nbytes = m * self.tp * K * x2d.element_size()
        if nbytes > self.max_gather_bytes:
            raise ValueError(f"gathered tensor of {nbytes} bytes exceeds the fused-TP workspace ({self.max_gather_bytes})")
        off = self.xg_off[parity]
        xg = self.buf[off: off + nbytes].view(torch.bfloat16).view(m * self.tp, K)
        xg[self.rank * m:(self.rank + 1) * m].copy_(x2d)
        if not self.push:
            xo = self.x_off[parity]
            self.buf[xo: xo + m * K * x2d.element_size()].view(torch.bfloat16).view(m, K).copy_(x2d)
        self._x_peer = [p + self.x_off[parity] for p in self.base_ptrs]
        return xg, [p + off for p in self.base_ptrs], [p + _FLAG_OFF for p in self.base_ptrs]

    def supports(self, rows_local: int, *dims: int) -> bool:
        return rows_local % 128 == 0 and all(d % 8 == 0 for d in dims)

    # ---- when does fusing pay?  (measured on B200: the fused kernels give up a few SMs to the transfer and run a
    # peer handshake, so they win when the NVLink time is a sizeable fraction of the GEMM time) -----------------
    GEMM_FLOPS = 1.35e15          # sustained tcgen05 bf16 rate of gemm_sm100 on Llama shapes
    LINK_BPS = 6.0e11             # ~600 GB/s effective inbound/outbound per GPU
    PER_CTA_BPS = 3.0e10          # measured: ~30 GB/s per comm CTA

    def _ratio(self, m_total: int, n: int, k: int, bytes_moved: float) -> float:
        return (bytes_moved / self.LINK_BPS) / max(2.0 * m_total * n * k / self.GEMM_FLOPS, 1e-9)

    # Measured on 2 / 8 B200s (tests/gpu_checks/check_fused_tp.py): fusing wins when the transfer is neither negligible
    # (ratio < lo: nothing to hide, the comm CTAs only take SMs away) nor dominant (ratio > hi: the GEMM is too short to
    # hide anything and NCCL's all-gather / reduce-scatter is as fast as the in-kernel transfer).
    def _in_band(self, ratio: float) -> bool:
        lo = float(os.environ.get("NXDT_FUSE_THRESHOLD", "0.30"))
        hi = float(os.environ.get("NXDT_FUSE_MAX_RATIO", "1.80"))
        return lo <= ratio <= hi

    def prefer_ag_gemm(self, rows_local: int, k: int, n: int) -> bool:
        return self._in_band(self._ratio(rows_local * self.tp, n, k, rows_local * (self.tp - 1) * k * 2))

    def prefer_ag_wgrad(self, rows_local: int, k: int, n: int) -> bool:
        return self._in_band(self._ratio(rows_local * self.tp, n, k, rows_local * (self.tp - 1) * k * 2))

    def prefer_gemm_rs(self, rows_total: int, k: int, n: int) -> bool:
        return self._in_band(self._ratio(rows_total, n, k, rows_total * n * 2 * (self.tp - 1) / self.tp))

    def _comm_ctas(self, m_total: int, n: int, k: int, bytes_in: float) -> int:
        """enough comm CTAs to finish the gather in ~half the GEMM time, bounded by what the link can take"""
        if self.num_comm_ctas_fixed:
            return self.num_comm_ctas_fixed
        # A comm CTA moves ~30 GB/s (every byte crosses its SM twice: bulk load + bulk store), so the gather takes
        # bytes_in / (nc * 30 GB/s) while the GEMM gets (148 - nc) SMs: pick nc that balances the two, capped by the link.
        t_gemm = 2.0 * m_total * n * k / self.GEMM_FLOPS
        best, best_t = 2, float("inf")
        for nc in range(2, 33, 2):
            t_comm = max(bytes_in / (nc * self.PER_CTA_BPS), bytes_in / self.LINK_BPS)
            t = max(t_comm, t_gemm * 148.0 / (148 - nc)) + 0.25 * min(t_comm, t_gemm)
            if t < best_t - 1e-9:
                best, best_t = nc, t
        return best

    # ---- ops ---------------------------------------------------------------------------------------
    def all_gather_gemm(self, x2d: torch.Tensor, w: torch.Tensor, trans_b: bool = True, return_gathered: bool = False):
        """``AG_rows(x2d) @ (w.T if trans_b else w)``; x2d [m, K] bf16 local shard."""
        m, K = x2d.shape
        N = w.shape[0] if trans_b else w.shape[1]
        epoch, par = self._next_epoch()
        x2d = x2d.contiguous()
        xg, xg_peer, flags_peer = self._gather_buffer(x2d, par)
        out = torch.empty(m * self.tp, N, dtype=x2d.dtype, device=x2d.device)
        hs_local, hs_peer = self._hs()
        ops._count()
        nc = self._comm_ctas(m * self.tp, N, K, m * (self.tp - 1) * K * 2)
        ops.C().ag_gemm(xg, x2d, w, out, out, xg_peer, flags_peer, hs_local, hs_peer, epoch, self.rank, self.tp,
                        0, trans_b, False, nc, self._x_peer, self.push, self._epoch_ptr)
        # NOTE: the gathered view lives in the double-buffered workspace: consume it before the second next fused call
        return (out, xg) if return_gathered else out

    def all_gather_wgrad(self, dy2d: torch.Tensor, x2d: torch.Tensor, out: torch.Tensor, accumulate: bool) -> torch.Tensor:
        """``out (+)= dy2d^T @ AG_rows(x2d)``; dy2d [M_total, N'], x2d [m, K'] local shard, out [N', K']."""
        m, K = x2d.shape
        epoch, par = self._next_epoch()
        x2d = x2d.contiguous()
        xg, xg_peer, flags_peer = self._gather_buffer(x2d, par)
        hs_local, hs_peer = self._hs()
        ops._count()
        nc = self._comm_ctas(dy2d.shape[1], K, m * self.tp, m * (self.tp - 1) * K * 2)
        ops.C().ag_gemm(xg, x2d, xg, out, dy2d.contiguous(), xg_peer, flags_peer, hs_local, hs_peer, epoch, self.rank,
                        self.tp, 1, True, accumulate, nc, self._x_peer, self.push, self._epoch_ptr)
        return out

    def all_gather(self, x2d: torch.Tensor) -> torch.Tensor:
        m, K = x2d.shape
        epoch, par = self._next_epoch()
        x2d = x2d.contiguous()
        xg, xg_peer, flags_peer = self._gather_buffer(x2d, par)
        hs_local, hs_peer = self._hs()
        ops._count()
        ops.C().ag_only(xg, x2d, xg_peer, flags_peer, hs_local, hs_peer, epoch, self.rank, self.tp, self.num_comm_ctas,
                        self._x_peer, self.push, self._epoch_ptr)
        return xg.clone()

    def gemm_reduce_scatter(self, a2d: torch.Tensor, w: torch.Tensor, trans_b: bool = True) -> torch.Tensor:
        """``RS_rows(a2d @ (w.T if trans_b else w))``; a2d [M_total, K_local]."""
        M, K = a2d.shape
        N = w.shape[0] if trans_b else w.shape[1]
        rows = M // self.tp
        need = self.tp * rows * N * 2
        if need > self.max_stage_bytes:
            raise ValueError(f"reduce-scatter staging of {need} bytes exceeds the workspace ({self.max_stage_bytes})")
        epoch, par = self._next_epoch()
        stage = [p + self.stage_off[par] for p in self.base_ptrs]
        cnt = [p + _CNT_OFF + par * _CNT_BYTES for p in self.base_ptrs]
        out = torch.empty(rows, N, dtype=a2d.dtype, device=a2d.device)
        hs_local, hs_peer = self._hs()
        ops._count()
        ops.C().gemm_rs(a2d.contiguous(), w, out, stage, cnt, hs_local, hs_peer, epoch, self.rank, self.tp, trans_b,
                        self._epoch_ptr)
        return out


class _FusedLayerAdapter:
    """What ``parallel.layers`` sees: falls back to ``None`` semantics per call when shapes are unsupported."""

    def __init__(self, ftp: FusedTP):
        self.ftp = ftp

    def ok(self, rows_local: int, *dims: int) -> bool:
        return self.ftp.supports(rows_local, *dims)

    def prefer_ag_gemm(self, rows_local, k, n):
        return self.ftp.prefer_ag_gemm(rows_local, k, n)

    def prefer_ag_wgrad(self, rows_local, k, n):
        return self.ftp.prefer_ag_wgrad(rows_local, k, n)

    def prefer_gemm_rs(self, rows_total, k, n):
        return self.ftp.prefer_gemm_rs(rows_total, k, n)

    def all_gather_gemm(self, x2d, w, trans_b=True, return_gathered=False):
        return self.ftp.all_gather_gemm(x2d, w, trans_b, return_gathered)

    def gemm_reduce_scatter(self, a2d, w, trans_b=True):
        return self.ftp.gemm_reduce_scatter(a2d, w, trans_b)

    def all_gather(self, x2d):
        return self.ftp.all_gather(x2d)

    def all_gather_wgrad(self, dy2d, x2d, out, accumulate):
        return self.ftp.all_gather_wgrad(dy2d, x2d, out, accumulate)


_FTP_CACHE = {}


def get_fused_tp(group=None, max_shard_bytes: int = 64 << 20, max_stage_bytes: int = 256 << 20) -> FusedTP:
    group = group if group is not None else ps.get_tensor_model_parallel_group()
    key = id(group)
    if key not in _FTP_CACHE:
        _FTP_CACHE[key] = FusedTP(group, max_shard_bytes, max_stage_bytes)
    return _FTP_CACHE[key]


def enable_fused_tp(model: nn.Module, seq_len: Optional[int] = None, micro_batch: int = 1) -> int:
    """Attach the fused kernels to every sequence-parallel Column/Row/GQA linear of ``model``.
    Returns the number of layers switched."""
    from .layers import ColumnParallelLinear, RowParallelLinear, GQAQKVColumnParallelLinear
    if not torch.cuda.is_available() or ops.C() is None or not hasattr(ops.C(), "ag_gemm"):
        raise RuntimeError("fused TP kernels need the sm_100a extension")
    tp = ps.get_tensor_model_parallel_size()
    if tp < 2:
        return 0
    mods = [m for m in model.modules() if isinstance(m, (ColumnParallelLinear, RowParallelLinear, GQAQKVColumnParallelLinear))
            and getattr(m, "sequence_parallel_enabled", False)]
    if not mods:
        return 0
    # size the workspace from the largest linear: shard = tokens_local x max(in features), staging = tp x
    # tokens_local x max(out features of a reduce-scatter)
    tokens_total = (seq_len or 8192) * micro_batch
    tokens_local = tokens_total // tp
    max_in = 0
    max_rs_out = 0
    for m in mods:
        if isinstance(m, RowParallelLinear):
            max_in = max(max_in, m.output_size)          # backward gathers dZ [tokens_local, out]
            max_rs_out = max(max_rs_out, m.output_size)  # forward reduce-scatters [tokens, out]
        else:
            max_in = max(max_in, m.input_size)
            max_rs_out = max(max_rs_out, m.input_size)   # dgrad reduce-scatters [tokens, in]
    ftp = get_fused_tp(mods[0].group, max_shard_bytes=tokens_local * max_in * 2,
                       max_stage_bytes=tokens_total * max_rs_out * 2)
    adapter = _FusedLayerAdapter(ftp)
    for m in mods:
        m.fused = adapter
    return len(mods)
