"""Pipeline-parallel engine: 1F1B and interleaved (virtual) 1F1B over point-to-point send/recv.

Capability parity: NxD ``NxDPPModel.run_train / run_eval`` (reference
lightning_modules/model/base.py:374-381): takes the whole per-DP-rank global batch, splits it into
``num_microbatches``, runs the schedule with activations/grads exchanged between adjacent stages and
returns the loss on the last stage (0 elsewhere; the caller all-reduces over the PP group).

P2P stays on NCCL (``batch_isend_irecv``) and is issued on a dedicated communication stream.  A round that only SENDS
never blocks the compute stream (the tensor is kept alive with ``record_stream``), so the transfer of microbatch *k*
overlaps the compute of microbatch *k+1*; the compute stream waits for a round only when it carries a receive, i.e. when
the next kernel really consumes the incoming tensor.
"""
from __future__ import annotations

import contextlib
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from ..parallel import state as ps


class P2PCommunicator:
    """send/recv of fixed-shape activation tensors between adjacent pipeline stages."""

    def __init__(self, shape: Tuple[int, ...], dtype: torch.dtype, device: torch.device):
        self.shape, self.dtype, self.device = tuple(shape), dtype, device
        self.group = ps.get_pipeline_model_parallel_group()
        self.prev = ps.get_pipeline_model_parallel_prev_rank()
        self.next = ps.get_pipeline_model_parallel_next_rank()
        self.first = ps.is_pipeline_first_stage()
        self.last = ps.is_pipeline_last_stage()
        self.stream = torch.cuda.Stream(device) if device.type == "cuda" else None

    def _buf(self):
        return torch.empty(self.shape, dtype=self.dtype, device=self.device)

    def _run(self, ops: List[dist.P2POp], has_recv: bool, sent: Tuple[torch.Tensor, ...] = ()):
        if not ops:
            return
        if self.stream is not None:
            # order after the producer on the compute stream and run on the comm stream
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                for r in dist.batch_isend_irecv(ops):
                    r.wait()                     # stream-level dependency for NCCL work, the host does not block
            for t in sent:
                t.record_stream(self.stream)     # the allocator must not recycle a buffer the transfer still reads
            if has_recv:
                # the very next kernel consumes the received tensor: the compute stream waits here.  Send-only rounds
                # return immediately and overlap with whatever the compute stream does next.
                torch.cuda.current_stream().wait_stream(self.stream)
        else:
            for r in dist.batch_isend_irecv(ops):
                r.wait()

    def exchange(self, send_next: Optional[torch.Tensor] = None, send_prev: Optional[torch.Tensor] = None,
                 recv_prev: bool = False, recv_next: bool = False):
        """One batched p2p round.  Returns (tensor_from_prev, tensor_from_next)."""
        # Ordering matters when prev == next (pipeline size 2 with virtual stages): messages between one
        # pair of ranks are matched in issue order, so receives must mirror the peer's send order
        # (gradients first, then activations).
        ops, from_prev, from_next, sent = [], None, None, []
        if send_prev is not None:
            sent.append(send_prev.contiguous())
            ops.append(dist.P2POp(dist.isend, sent[-1], self.prev, self.group))
        if send_next is not None:
            sent.append(send_next.contiguous())
            ops.append(dist.P2POp(dist.isend, sent[-1], self.next, self.group))
        if recv_next:
            from_next = self._buf()
            ops.append(dist.P2POp(dist.irecv, from_next, self.next, self.group))
        if recv_prev:
            from_prev = self._buf()
            ops.append(dist.P2POp(dist.irecv, from_prev, self.prev, self.group))
        self._run(ops, recv_prev or recv_next, tuple(sent))
        return from_prev, from_next

    # convenience wrappers that respect the first/last stage
    def recv_forward(self):
        return None if self.first else self.exchange(recv_prev=True)[0]

    def recv_backward(self):
        return None if self.last else self.exchange(recv_next=True)[1]

    def send_forward(self, out):
        if not self.last:
            self.exchange(send_next=out)

    def send_backward(self, grad):
        if not self.first:
            self.exchange(send_prev=grad)

    def send_forward_recv_backward(self, out):
        return None if self.last else self.exchange(send_next=out, recv_next=True)[1]

    def send_backward_recv_forward(self, grad):
        return None if self.first else self.exchange(send_prev=grad, recv_prev=True)[0]


class PipelineSchedule:
    """Runs microbatches through this rank's model chunk(s).

    ``forward_fn(chunk_idx, microbatch_dict, hidden_in) -> hidden_out | loss`` is supplied by the model
    wrapper; the schedule is agnostic of the model.
    """

    def __init__(self, forward_fn: Callable, num_chunks: int, comm: P2PCommunicator, num_microbatches: int):
        self.forward_fn, self.num_chunks, self.comm, self.n_mb = forward_fn, num_chunks, comm, num_microbatches
        self.pp = ps.get_pipeline_model_parallel_size()
        self.rank = ps.get_pipeline_model_parallel_rank()

    # ---- single steps ------------------------------------------------------------------------------
    def _is_first_chunk(self, chunk: int) -> bool:
        return self.comm.first and chunk == 0

    def _is_last_chunk(self, chunk: int) -> bool:
        return self.comm.last and chunk == self.num_chunks - 1

    def _forward_step(self, chunk: int, mb: Dict[str, Any], inp: Optional[torch.Tensor], train: bool):
        if inp is not None and train:
            inp.requires_grad_(True)
        ctx = contextlib.nullcontext() if train else torch.no_grad()
        with ctx:
            out = self.forward_fn(chunk, mb, inp)
        if self._is_last_chunk(chunk):
            out = out.float() / self.n_mb
        return out

    @staticmethod
    def _backward_step(inp: Optional[torch.Tensor], out: torch.Tensor, out_grad: Optional[torch.Tensor]):
        if out_grad is None:
            out.backward()
        else:
            torch.autograd.backward(out, out_grad)
        g = None
        if inp is not None:
            g = inp.grad
            inp.grad = None
        return g

    # ---- non-interleaved 1F1B --------------------------------------------------------------------------
    def run_1f1b(self, microbatches: List[Dict[str, Any]], train: bool = True) -> torch.Tensor:
        comm, n = self.comm, len(microbatches)
        dev = comm.device
        loss_acc = torch.zeros((), dtype=torch.float32, device=dev)
        if not train:
            for mb in microbatches:
                inp = comm.recv_forward()
                out = self._forward_step(0, mb, inp, False)
                if comm.last:
                    loss_acc += out.detach()
                else:
                    comm.send_forward(out)
            return loss_acc
        warm = min(self.pp - self.rank - 1, n)
        remaining = n - warm
        ins, outs = [], []
        it = iter(microbatches)
        for _ in range(warm):
            inp = comm.recv_forward()
            out = self._forward_step(0, next(it), inp, True)
            comm.send_forward(out)
            ins.append(inp); outs.append(out)
        inp = comm.recv_forward() if remaining > 0 else None
        for i in range(remaining):
            out = self._forward_step(0, next(it), inp, True)
            if comm.last:
                loss_acc += out.detach()
            out_grad = comm.send_forward_recv_backward(out)
            ins.append(inp); outs.append(out)
            i0, o0 = ins.pop(0), outs.pop(0)
            g = self._backward_step(i0, o0, out_grad)
            if i == remaining - 1:
                comm.send_backward(g)
            else:
                inp = comm.send_backward_recv_forward(g)
        for _ in range(warm):
            i0, o0 = ins.pop(0), outs.pop(0)
            out_grad = comm.recv_backward()
            g = self._backward_step(i0, o0, out_grad)
            comm.send_backward(g)
        return loss_acc

    # ---- interleaved (virtual pipeline) 1F1B ------------------------------------------------------------
    def _chunk_id(self, k: int, forward: bool) -> int:
        g = k % (self.pp * self.num_chunks)
        c = g // self.pp
        return c if forward else self.num_chunks - c - 1

    def run_interleaved(self, microbatches: List[Dict[str, Any]], train: bool = True) -> torch.Tensor:
        comm, V, pp, rank = self.comm, self.num_chunks, self.pp, self.rank
        n = len(microbatches)
        if n % pp != 0:
            raise ValueError(f"interleaved schedule needs num_microbatches ({n}) % pipeline size ({pp}) == 0")
        dev = comm.device
        loss_acc = torch.zeros((), dtype=torch.float32, device=dev)
        total = n * V
        ins = [[] for _ in range(V)]
        outs = [[] for _ in range(V)]
        out_grads = [[] for _ in range(V)]
        fwd_mb_counter = [0] * V

        def mb_for(k: int) -> Dict[str, Any]:
            # virtual step k processes microbatch: group of pp consecutive microbatches per chunk visit
            grp = k // (pp * V)
            within = k % pp
            return microbatches[grp * pp + within]

        if not train:
            warm = total
        else:
            warm = min((pp - rank - 1) * 2 + (V - 1) * pp, total)
        remaining = total - warm

        def fwd(k: int):
            c = self._chunk_id(k, True)
            inp = ins[c][-1] if len(ins[c]) > len(outs[c]) else None
            if self._is_first_chunk(c):
                if len(ins[c]) == len(outs[c]):
                    ins[c].append(None)
                inp = None
            out = self._forward_step(c, mb_for(k), inp, train)
            outs[c].append(out)
            if self._is_last_chunk(c):
                nonlocal loss_acc
                loss_acc += out.detach()
            if not train:
                ins[c].pop(); outs[c].pop()
            return out

        def bwd(k: int):
            c = self._chunk_id(k, False)
            if self._is_last_chunk(c) and len(out_grads[c]) == 0:
                out_grads[c].append(None)
            i0, o0, g0 = ins[c].pop(0), outs[c].pop(0), out_grads[c].pop(0)
            return self._backward_step(i0, o0, g0)

        # first receive
        ins[0].append(comm.recv_forward() if not self._is_first_chunk(0) else None)
        for k in range(warm):
            out = fwd(k)
            nxt = self._chunk_id(k + 1, True) if k + 1 < total else None
            recv_prev = nxt is not None and not self._is_first_chunk(nxt)
            send = None if self._is_last_chunk(self._chunk_id(k, True)) else out
            if k == warm - 1 and train and remaining > 0:
                # overlap the first backward receive with the last warm-up send
                c_b = self._chunk_id(0, False)
                recv_next = not self._is_last_chunk(c_b)
                fp, fn = comm.exchange(send_next=send, recv_prev=recv_prev, recv_next=recv_next)
                if recv_next:
                    out_grads[c_b].append(fn)
            else:
                fp, _ = comm.exchange(send_next=send, recv_prev=recv_prev)
            if nxt is not None:
                ins[nxt].append(fp if recv_prev else None)
        if not train:
            return loss_acc
        for i in range(remaining):
            kf, kb = warm + i, i
            out = fwd(kf)
            g = bwd(kb)
            cf, cb = self._chunk_id(kf, True), self._chunk_id(kb, False)
            send_next = None if self._is_last_chunk(cf) else out
            send_prev = None if self._is_first_chunk(cb) else g
            nf = self._chunk_id(kf + 1, True) if kf + 1 < total else None
            recv_prev = nf is not None and not self._is_first_chunk(nf)
            nb = self._chunk_id(kb + 1, False) if kb + 1 < total else None
            recv_next = nb is not None and not self._is_last_chunk(nb)
            fp, fn = comm.exchange(send_next=send_next, send_prev=send_prev, recv_prev=recv_prev, recv_next=recv_next)
            if nf is not None:
                ins[nf].append(fp if recv_prev else None)
            if nb is not None and recv_next:
                out_grads[nb].append(fn)
        for kb in range(remaining, total):
            if remaining == 0 and kb == 0:
                c_b = self._chunk_id(0, False)
                if not self._is_last_chunk(c_b):
                    out_grads[c_b].append(comm.exchange(recv_next=True)[1])
            g = bwd(kb)
            cb = self._chunk_id(kb, False)
            send_prev = None if self._is_first_chunk(cb) else g
            nb = self._chunk_id(kb + 1, False) if kb + 1 < total else None
            recv_next = nb is not None and not self._is_last_chunk(nb)
            _, fn = comm.exchange(send_prev=send_prev, recv_next=recv_next)
            if nb is not None and recv_next:
                out_grads[nb].append(fn)
        return loss_acc

    def run(self, microbatches: List[Dict[str, Any]], train: bool = True) -> torch.Tensor:
        if self.num_chunks > 1:
            return self.run_interleaved(microbatches, train)
        return self.run_1f1b(microbatches, train)
