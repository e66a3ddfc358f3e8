"""Sharded checkpoint save / load in the reference's on-disk layout.

Layout (reference nlp_overrides.py:535-640, docs/tutorials/checkpoint_conversion.rst:137-141):

    <dir>/<tag>/
        done , checkpoint                                     marker files
        model/dp_rank_00_tp_rank_TT_pp_rank_PP.pt             (DP rank 0 saves the model)
        model/dp_rank_00_tp_rank_TT_pp_rank_PP.pt.tensors/tensor_<id>.pt      ("xser": one file per tensor)
        model/dp_rank_00_tp_rank_TT_pp_rank_PP.pt.info.pt     id -> dtype/shape table
        optim/dp_rank_DD_tp_rank_TT_pp_rank_PP.pt(+.tensors)  every DP rank saves its ZeRO-1 shard
        user_content.pt                                       loops / schedulers / callbacks / hparams

Capability parity with ``nxd.save_checkpoint / load_checkpoint / finalize_checkpoint`` and
``g_iostate.submit_remove``.  The "xser" tensor-per-file mode streams tensors one at
a time through pinned host staging so host RAM stays bounded; ``async_save`` hands the staged buffers to
the native writer thread-pool (``checkpoint/csrc/ckpt_io.cpp``) instead of forking a process, which
avoids the fork-doubles-host-memory problem the reference documents (known_issues.rst:53-81).
"""
from __future__ import annotations

import logging
import os
import shutil
import struct
import threading
import time
from concurrent.futures import ThreadPoolExecutor, Future
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from ..parallel import state as ps
from . import storage

log = logging.getLogger(__name__)

_MAGIC = b"NXDTB200"
_DTYPES = [torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int8, torch.uint8, torch.int16,
           torch.int32, torch.int64, torch.bool, torch.float8_e4m3fn, torch.float8_e5m2]
_DTYPE_CODE = {d: i for i, d in enumerate(_DTYPES)}


class TensorReference:
    """Placeholder stored in the main .pt file in xser mode."""
    __slots__ = ("tid",)

    def __init__(self, tid: int):
        self.tid = tid

    def __repr__(self):
        return f"TensorReference({self.tid})"


EXTRA_FILES_KEY = "_extra_files"      # user_content[EXTRA_FILES_KEY] = {file name: text}: written next to user_content.pt

def _native_io():
    try:
        from . import native_io
        return native_io.get()
    except Exception:
        return None


# ---- raw tensor files --------------------------------------------------------------------------------
def _tensor_header(t: torch.Tensor) -> bytes:
    shape = list(t.shape)
    h = _MAGIC + struct.pack("<II", _DTYPE_CODE[t.dtype], len(shape)) + struct.pack(f"<{len(shape)}q", *shape)
    pad = (-len(h)) % 64
    return h + b"\0" * pad


def _write_tensor_file(path: str, t: torch.Tensor) -> None:
    t = t.detach().contiguous()
    if t.device.type != "cpu":
        t = t.cpu()
    hdr = _tensor_header(t)
    nio = _native_io()
    raw = t.view(torch.uint8).reshape(-1) if t.numel() > 0 else torch.empty(0, dtype=torch.uint8)
    if nio is not None:
        nio.write_file(path, hdr, raw)
        return
    with open(path, "wb") as f:
        f.write(hdr)
        if raw.numel():
            f.write(raw.numpy().tobytes() if raw.numel() < (1 << 20) else memoryview(raw.numpy()))


def _read_tensor_file(path: str) -> torch.Tensor:
    with open(path, "rb") as f:
        head = f.read(16)
        if head[:8] != _MAGIC:
            # a torch.save'd per-tensor file (reference / torch-xla xser layout)
            f.seek(0)
            return torch.load(f, map_location="cpu", weights_only=False)
        code, nd = struct.unpack("<II", head[8:16])
        shape = struct.unpack(f"<{nd}q", f.read(8 * nd)) if nd else ()
        hdr_len = 16 + 8 * nd
        hdr_len += (-hdr_len) % 64
        dtype = _DTYPES[code]
        n = 1
        for s in shape:
            n *= s
        out = torch.empty(n * torch.empty((), dtype=dtype).element_size(), dtype=torch.uint8)
        nio = _native_io()
        if nio is not None and n > 0:
            nio.read_file(path, hdr_len, out)
        elif n > 0:
            f.seek(hdr_len)
            f.readinto(memoryview(out.numpy()))
        return out.view(dtype).reshape(shape)


# ---- async IO state ----------------------------------------------------------------------------------
class _IOState:
    """Background writer + ordered removal queue (``g_iostate`` equivalent)."""

    def __init__(self, workers: int = 4):
        self.pool = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="ckpt-io")
        self.pending: List[Future] = []
        self.lock = threading.Lock()

    def submit(self, fn, *a, **kw) -> Future:
        fut = self.pool.submit(fn, *a, **kw)
        with self.lock:
            self.pending.append(fut)
        return fut

    def wait_all(self):
        with self.lock:
            futs, self.pending = self.pending, []
        for f in futs:
            f.result()

    def submit_remove(self, path: str, async_remove: bool = True):
        def rm():
            if os.path.isdir(path):
                shutil.rmtree(path, ignore_errors=True)
            elif os.path.exists(path):
                os.remove(path)
        if async_remove:
            # removal must not overtake an in-flight save of the same tag
            prior = list(self.pending)

            def ordered():
                for f in prior:
                    try:
                        f.result()
                    except Exception:
                        pass
                rm()
            self.submit(ordered)
        else:
            self.wait_all()
            rm()


g_iostate = _IOState()


# ---- helpers -------------------------------------------------------------------------------------------
def _is_remote(path: str) -> bool:
    return storage.is_remote(path)


def shard_file_name(dp_rank: int, tp_rank: int, pp_rank: int) -> str:
    return f"dp_rank_{dp_rank:02d}_tp_rank_{tp_rank:02d}_pp_rank_{pp_rank:02d}.pt"


def _flatten_tensors(obj, table: List[torch.Tensor]):
    """Replace every tensor in a nested container by a TensorReference (dedup by storage identity)."""
    if isinstance(obj, torch.Tensor):
        table.append(obj)
        return TensorReference(len(table) - 1)
    if isinstance(obj, dict):
        return {k: _flatten_tensors(v, table) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        r = [_flatten_tensors(v, table) for v in obj]
        return type(obj)(r) if not hasattr(obj, "_fields") else type(obj)(*r)
    return obj


def _restore_tensors(obj, loader):
    if isinstance(obj, TensorReference):
        return loader(obj.tid)
    if isinstance(obj, dict):
        return {k: _restore_tensors(v, loader) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        r = [_restore_tensors(v, loader) for v in obj]
        return type(obj)(r) if not hasattr(obj, "_fields") else type(obj)(*r)
    return obj


def _stage_to_host(t: torch.Tensor) -> torch.Tensor:
    """Snapshot a tensor into (pinned when possible) host memory."""
    if t.device.type == "cpu":
        return t.detach().clone()
    host = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
    host.copy_(t.detach(), non_blocking=True)
    return host


def _save_object(obj: Any, path: str, use_xser: bool, async_save: bool) -> None:
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if not use_xser:
        table: List[torch.Tensor] = []
        staged = _restore_tensors(_flatten_tensors(obj, table), lambda i: _stage_to_host(table[i]))
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()
        if async_save:
            g_iostate.submit(torch.save, staged, path)
        else:
            torch.save(staged, path)
        return
    table = []
    skeleton = _flatten_tensors(obj, table)
    tdir = path + ".tensors"
    os.makedirs(tdir, exist_ok=True)
    info = {i: (str(t.dtype), tuple(t.shape)) for i, t in enumerate(table)}
    torch.save(skeleton, path)
    torch.save(info, path + ".info.pt")
    # stream tensor by tensor: bounded host memory in sync mode, full snapshot in async mode
    for i, t in enumerate(table):
        host = _stage_to_host(t)
        fpath = os.path.join(tdir, f"tensor_{i}.pt")
        if async_save:
            if t.device.type != "cpu":
                torch.cuda.current_stream().synchronize()
            g_iostate.submit(_write_tensor_file, fpath, host)
        else:
            if t.device.type != "cpu":
                torch.cuda.current_stream().synchronize()
            _write_tensor_file(fpath, host)


def _load_object(path: str, map_location="cpu") -> Any:
    skeleton = torch.load(path, map_location="cpu", weights_only=False)
    tdir = path + ".tensors"
    if os.path.isdir(tdir):
        def loader(i):
            t = _read_tensor_file(os.path.join(tdir, f"tensor_{i}.pt"))
            return t if map_location in ("cpu", None) else t.to(map_location)
        return _restore_tensors(skeleton, loader)
    return skeleton


# ---- public API ------------------------------------------------------------------------------------------
def save_checkpoint(checkpoint_dir: str, tag: str, model: Optional[Dict[str, Any]] = None,
                    optimizer: Optional[Dict[str, Any]] = None, user_content: Optional[Dict[str, Any]] = None,
                    use_xser: bool = True, async_save: bool = False, zero1_optimizer: bool = True,
                    num_kept_ckpts: Optional[int] = None, _publish: bool = True) -> str:
    """Every rank calls this.  ``model`` / ``optimizer`` are (already sharded) state dicts.  ``checkpoint_dir`` may be
    an object-store URL (``s3://bucket/prefix``): shards are written to node-local staging and uploaded, the ``done``
    marker last (see ``checkpoint/storage.py``)."""
    if _is_remote(checkpoint_dir):
        return _save_checkpoint_remote(checkpoint_dir, tag, model, optimizer, user_content, use_xser, async_save,
                                       zero1_optimizer)
    root = os.path.join(checkpoint_dir, tag)
    dp_rank = ps.get_data_parallel_rank() if ps.get_context_model_parallel_size() == 1 else \
        ps.group_rank(ps.get_data_context_parallel_group())
    tp_rank, pp_rank = ps.get_tensor_model_parallel_rank(), ps.get_pipeline_model_parallel_rank()
    rank = ps.get_rank()
    if rank == 0:
        os.makedirs(os.path.join(root, "model"), exist_ok=True)
        os.makedirs(os.path.join(root, "optim"), exist_ok=True)
        for marker in ("done",):
            mp_ = os.path.join(root, marker)
            if os.path.exists(mp_):
                os.remove(mp_)
        # completion markers of an earlier (possibly torn) save of the same tag must not count for this one
        shutil.rmtree(os.path.join(root, ".ranks"), ignore_errors=True)
    ps.barrier()
    if model is not None and dp_rank == 0:
        _save_object(model, os.path.join(root, "model", shard_file_name(0, tp_rank, pp_rank)), use_xser, async_save)
    if optimizer is not None and (zero1_optimizer or dp_rank == 0):
        _save_object(optimizer, os.path.join(root, "optim", shard_file_name(dp_rank, tp_rank, pp_rank)),
                     use_xser, async_save)
    if user_content is not None and rank == 0:
        # small side files (e.g. the PEFT adapter_config.json) travel as {"file name": text} under a reserved key
        extra = user_content.pop(EXTRA_FILES_KEY, None) if isinstance(user_content, dict) else None
        torch.save(user_content, os.path.join(root, "user_content.pt"))
        for name, text in (extra or {}).items():
            with open(os.path.join(root, os.path.basename(name)), "w") as f:
                f.write(text)
        if extra and not _publish:          # staging for an object store: tell the uploader which side files exist
            with open(os.path.join(root, ".extra_files"), "w") as f:
                f.write("\n".join(os.path.basename(n) for n in extra) + "\n")

    def finish():
        # marker files: written once every rank's writes are durable
        if rank == 0:
            with open(os.path.join(root, "checkpoint"), "w") as f:
                f.write(tag + "\n")
            with open(os.path.join(root, "done"), "w") as f:
                f.write(str(time.time()) + "\n")

    if not _publish:
        return root          # the caller (object-store path) publishes the markers itself
    if async_save:
        # per-rank completion markers let rank 0 publish "done" without a collective on the IO thread
        mdir = os.path.join(root, ".ranks")
        os.makedirs(mdir, exist_ok=True)
        world = ps.get_world_size()
        my_futs = list(g_iostate.pending)

        def mark():
            for f in my_futs:
                f.result()
            open(os.path.join(mdir, f"{rank}.ok"), "w").close()
            if rank == 0:
                deadline = time.time() + 3600
                while len([n for n in os.listdir(mdir) if n.endswith(".ok")]) < world and time.time() < deadline:
                    time.sleep(0.05)
                if len([n for n in os.listdir(mdir) if n.endswith(".ok")]) < world:
                    log.error("checkpoint %s: not every rank finished writing within the deadline; leaving it WITHOUT a "
                              "'done' marker so that auto-resume skips it", root)
                    return
                finish()
        g_iostate.submit(mark)
    else:
        ps.barrier()
        finish()
        ps.barrier()
    return root


def _my_shard_names() -> Tuple[str, str]:
    dp_rank = ps.get_data_parallel_rank() if ps.get_context_model_parallel_size() == 1 else \
        ps.group_rank(ps.get_data_context_parallel_group())
    tp_rank, pp_rank = ps.get_tensor_model_parallel_rank(), ps.get_pipeline_model_parallel_rank()
    return shard_file_name(0, tp_rank, pp_rank), shard_file_name(dp_rank, tp_rank, pp_rank)


def _save_checkpoint_remote(url_dir, tag, model, optimizer, user_content, use_xser, async_save, zero1_optimizer) -> str:
    store = storage.get_store(url_dir)
    local_dir = storage.staging_dir(url_dir)
    root_url = storage.join_url(url_dir, tag)
    rank, world = ps.get_rank(), ps.get_world_size()
    dp_rank = ps.get_data_parallel_rank() if ps.get_context_model_parallel_size() == 1 else \
        ps.group_rank(ps.get_data_context_parallel_group())
    if rank == 0:
        store.remove_prefix(root_url)          # a re-used tag must not keep a stale 'done'
    ps.barrier()
    local_root = save_checkpoint(local_dir, tag, model, optimizer, user_content, use_xser, async_save, zero1_optimizer,
                                 _publish=False)
    model_name, optim_name = _my_shard_names()
    mine = []
    if model is not None and dp_rank == 0:
        mine.append("model/" + model_name)
    if optimizer is not None and (zero1_optimizer or dp_rank == 0):
        mine.append("optim/" + optim_name)
    pending = list(g_iostate.pending)

    def publish():
        for f in pending:
            f.result()
        store.upload_tree(local_root, root_url, only=lambda rel: any(rel.startswith(m) for m in mine))
        marker = os.path.join(local_root, f".uploaded_{rank}")
        open(marker, "w").close()
        store.upload_file(marker, storage.join_url(root_url, ".ranks", f"{rank}.ok"))
        os.remove(marker)
        for dirpath, _, files in os.walk(local_root):          # this rank's staged copies are no longer needed
            for f in files:
                rel = os.path.relpath(os.path.join(dirpath, f), local_root).replace(os.sep, "/")
                if any(rel.startswith(m) for m in mine):
                    os.remove(os.path.join(dirpath, f))
        if rank == 0:
            deadline = time.time() + 3600
            complete = False
            while time.time() < deadline and not complete:
                complete = len([u for u in store.list(storage.join_url(root_url, ".ranks")) if u.endswith(".ok")]) >= world
                if not complete:
                    time.sleep(0.05)
            if not complete:
                log.error("checkpoint %s: not every rank finished uploading within the deadline; not publishing 'done'", root_url)
                return
            with open(os.path.join(local_root, "checkpoint"), "w") as f:
                f.write(tag + "\n")
            with open(os.path.join(local_root, "done"), "w") as f:
                f.write(str(time.time()) + "\n")
            extra_names = []
            if os.path.exists(os.path.join(local_root, ".extra_files")):
                with open(os.path.join(local_root, ".extra_files")) as f:
                    extra_names = [l.strip() for l in f if l.strip()]
                os.remove(os.path.join(local_root, ".extra_files"))
            for name in (*extra_names, "user_content.pt", "checkpoint", "done"):       # 'done' strictly last
                lp = os.path.join(local_root, name)
                if os.path.exists(lp):
                    store.upload_file(lp, storage.join_url(root_url, name))
                    os.remove(lp)

    if async_save:
        g_iostate.submit(publish)
    else:
        publish()
        ps.barrier()
    return root_url


def _load_checkpoint_remote(url_dir, tag, model, optimizer, strict_done, map_location):
    store = storage.get_store(url_dir)
    root_url = storage.join_url(url_dir, tag)
    if strict_done and not store.exists(storage.join_url(root_url, "done")):
        raise FileNotFoundError(f"checkpoint {root_url} has no 'done' marker (incomplete save?)")
    local_dir = os.path.join(storage.staging_dir(url_dir), f"load_rank{ps.get_rank()}")
    local_root = os.path.join(local_dir, tag)
    shutil.rmtree(local_root, ignore_errors=True)
    os.makedirs(local_root, exist_ok=True)
    model_name, optim_name = _my_shard_names()
    wanted = []
    if model:
        wanted.append("model/" + model_name)
    if optimizer:
        have = store.list(storage.join_url(root_url, "optim"))
        prefix = storage.join_url(root_url, "optim", optim_name)
        wanted.append("optim/" + (optim_name if any(u.startswith(prefix) for u in have) else model_name))
    for w in wanted:
        sub = w.split("/")[0]
        for u in store.list(storage.join_url(root_url, sub)):
            if u.startswith(storage.join_url(root_url, w)):
                rel = u[len(root_url) + 1:]
                lp = os.path.join(local_root, *rel.split("/"))
                os.makedirs(os.path.dirname(lp), exist_ok=True)
                store.download_file(u, lp)
    for name in ("user_content.pt", "done"):
        u = storage.join_url(root_url, name)
        if store.exists(u):
            store.download_file(u, os.path.join(local_root, name))
    try:
        return load_checkpoint(local_dir, tag, model, optimizer, strict_done, map_location)
    finally:
        shutil.rmtree(local_root, ignore_errors=True)


def finalize_checkpoint() -> None:
    """Block until all asynchronous writes / removals are done (call at teardown)."""
    g_iostate.wait_all()
    ps.barrier()


def has_checkpoint(checkpoint_dir: str, tag: str) -> bool:
    if _is_remote(checkpoint_dir):
        return storage.get_store(checkpoint_dir).exists(storage.join_url(checkpoint_dir, tag, "done"))
    return os.path.exists(os.path.join(checkpoint_dir, tag, "done"))


def load_checkpoint(checkpoint_dir: str, tag: str, model: bool = True, optimizer: bool = True,
                    strict_done: bool = True, map_location="cpu") -> Tuple[Optional[dict], Optional[dict], Optional[dict]]:
    """Returns (model_state, optimizer_state, user_content) for THIS rank's shard."""
    if _is_remote(checkpoint_dir):
        return _load_checkpoint_remote(checkpoint_dir, tag, model, optimizer, strict_done, map_location)
    root = os.path.join(checkpoint_dir, tag)
    if strict_done and not os.path.exists(os.path.join(root, "done")):
        raise FileNotFoundError(f"checkpoint {root} has no 'done' marker (incomplete save?)")
    dp_rank = ps.get_data_parallel_rank() if ps.get_context_model_parallel_size() == 1 else \
        ps.group_rank(ps.get_data_context_parallel_group())
    tp_rank, pp_rank = ps.get_tensor_model_parallel_rank(), ps.get_pipeline_model_parallel_rank()
    m = o = None
    if model:
        mpath = os.path.join(root, "model", shard_file_name(0, tp_rank, pp_rank))
        if _broadcast_load_enabled():
            # every data-parallel replica needs the same model shard: ONE rank per (tp, pp) reads it and the others
            # receive it over the data-parallel group, instead of dp_size readers hammering the same files
            # (reference docs/general/features.rst: "load only on 1 data-parallel rank and broadcast")
            m = _broadcast_over_dp(_load_object(mpath, map_location) if dp_rank == 0 else None)
        else:
            m = _load_object(mpath, map_location)
    if optimizer:
        p = os.path.join(root, "optim", shard_file_name(dp_rank, tp_rank, pp_rank))
        if not os.path.exists(p):
            p = os.path.join(root, "optim", shard_file_name(0, tp_rank, pp_rank))
        if os.path.exists(p):
            o = _load_object(p, map_location)
    uc_path = os.path.join(root, "user_content.pt")
    uc = torch.load(uc_path, map_location="cpu", weights_only=False) if os.path.exists(uc_path) else None
    return m, o, uc


def _dp_load_group():
    return ps.get_data_parallel_group() if ps.get_context_model_parallel_size() == 1 else ps.get_data_context_parallel_group()


def _broadcast_load_enabled() -> bool:
    """On by default whenever there is more than one data-parallel replica (``NXDT_CKPT_LOAD_BROADCAST=0`` turns it off)."""
    if os.environ.get("NXDT_CKPT_LOAD_BROADCAST", "1") == "0" or not dist.is_initialized():
        return False
    return ps.group_size(_dp_load_group()) > 1


def _broadcast_over_dp(obj):
    """``obj`` (a nested dict / list of tensors and plain values) is given on data-parallel rank 0 and ``None`` elsewhere;
    returns it on every rank of the data-parallel group.  The structure travels as one pickled skeleton, the tensors
    one by one (through the GPU when the group's backend is NCCL)."""
    group = _dp_load_group()
    src = dist.get_global_rank(group, 0)
    is_src = dist.get_rank() == src
    tensors: List[torch.Tensor] = []

    def strip(x):
        if torch.is_tensor(x):
            tensors.append(x)
            return ("__tensor__", len(tensors) - 1, tuple(x.shape), x.dtype)
        if isinstance(x, dict):
            return {k: strip(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return type(x)(strip(v) for v in x)
        return x

    box = [strip(obj) if is_src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    skeleton = box[0]
    use_cuda = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    received: Dict[int, torch.Tensor] = {}

    def fill(x):
        if isinstance(x, tuple) and len(x) == 4 and x[0] == "__tensor__":
            _, idx, shape, dtype = x
            if idx not in received:
                t = tensors[idx].to(dev).contiguous() if is_src else torch.empty(shape, dtype=dtype, device=dev)
                dist.broadcast(t, src=src, group=group)
                # hand back host tensors, like a read from disk would: the GPU copy is only the NCCL staging buffer
                received[idx] = tensors[idx] if is_src else (t.cpu() if use_cuda else t)
            return received[idx]
        if isinstance(x, dict):
            return {k: fill(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return type(x)(fill(v) for v in x)
        return x

    return fill(skeleton)


def remove_checkpoint(path: str, async_remove: bool = True) -> None:
    if _is_remote(path):
        if ps.get_rank() == 0:
            store = storage.get_store(path)
            if async_remove:
                g_iostate.submit(store.remove_prefix, path)
            else:
                store.remove_prefix(path)
        return
    if ps.get_rank() == 0:
        g_iostate.submit_remove(path, async_remove)
