"""Megatron-format pre-training data: mmap ``.bin/.idx`` indexed datasets, GPT sample index maps, blending,
resumable batch samplers, tokenizers.

Capability parity: reference megatron/data_module.py:25-356 and datasets/gpt_dataset_patch.py:53-570 (+ the NeMo
indexed-dataset / blendable-dataset / sampler pieces they import):
  * the on-disk format is Megatron's ``MMIDIDX`` index (+ raw token ``.bin``), readable and writable here;
  * ``doc_idx`` / ``sample_idx`` / ``shuffle_idx`` are built on (global) rank 0, cached as ``.npy`` next to the data
    (or under ``index_mapping_dir``) and mmapped by everybody after a barrier; ``sample_idx`` comes from the C++
    helper (``data/csrc/helpers.cpp``) with a numpy fallback used as the test oracle;
  * per-sample ``tokens / labels / loss_mask / position_ids`` with the EOD reset options.
"""
from __future__ import annotations

import hashlib
import logging
import os
import struct
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from ..config import Config
from ..parallel import state as ps
from .base import BaseDataModule, ResumableBatchSampler, default_collate
from . import helpers_build

log = logging.getLogger("nxdt")

_HDR_MAGIC = b"MMIDIDX\x00\x00"
_DTYPES = {1: np.uint8, 2: np.int8, 3: np.int16, 4: np.int32, 5: np.int64, 6: np.float32, 7: np.float64, 8: np.uint16}
_DTYPE_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


# =====================================================================================================
# indexed dataset
# =====================================================================================================
class MMapIndexedDataset:
    """Reader for ``<prefix>.idx`` + ``<prefix>.bin``."""

    def __init__(self, prefix: str):
        self.prefix = prefix
        with open(prefix + ".idx", "rb") as f:
            magic = f.read(9)
            if magic != _HDR_MAGIC:
                raise ValueError(f"{prefix}.idx is not a Megatron mmap index (bad magic)")
            (version,) = struct.unpack("<Q", f.read(8))
            assert version == 1, f"unsupported index version {version}"
            (code,) = struct.unpack("<B", f.read(1))
            self.dtype = np.dtype(_DTYPES[code])
            (self._len,) = struct.unpack("<Q", f.read(8))
            (self._doc_count,) = struct.unpack("<Q", f.read(8))
            offset = f.tell()
        self._idx_mmap = np.memmap(prefix + ".idx", mode="r", order="C")
        self.sizes = np.frombuffer(self._idx_mmap, dtype=np.int32, count=self._len, offset=offset)
        self.pointers = np.frombuffer(self._idx_mmap, dtype=np.int64, count=self._len, offset=offset + self.sizes.nbytes)
        self.doc_idx = np.frombuffer(self._idx_mmap, dtype=np.int64, count=self._doc_count,
                                     offset=offset + self.sizes.nbytes + self.pointers.nbytes)
        self._bin = np.memmap(prefix + ".bin", mode="r", order="C")

    def __len__(self):
        return self._len

    def get(self, idx: int, offset: int = 0, length: Optional[int] = None) -> np.ndarray:
        size = int(self.sizes[idx])
        length = size - offset if length is None else length
        ptr = int(self.pointers[idx]) + offset * self.dtype.itemsize
        return np.frombuffer(self._bin, dtype=self.dtype, count=length, offset=ptr)

    def __getitem__(self, idx: int) -> np.ndarray:
        return self.get(idx)

    @staticmethod
    def exists(prefix: str) -> bool:
        return os.path.exists(prefix + ".idx") and os.path.exists(prefix + ".bin")


class MMapIndexedDatasetBuilder:
    """Writer (used by preprocessing tools and the tests)."""

    def __init__(self, prefix: str, dtype=np.int32):
        self.prefix, self.dtype = prefix, np.dtype(dtype)
        self._bin = open(prefix + ".bin", "wb")
        self.sizes: List[int] = []
        self.doc_idx: List[int] = [0]

    def add_item(self, tokens: Sequence[int]):
        arr = np.asarray(tokens, dtype=self.dtype)
        self._bin.write(arr.tobytes(order="C"))
        self.sizes.append(arr.size)

    def end_document(self):
        self.doc_idx.append(len(self.sizes))

    def finalize(self):
        self._bin.close()
        sizes = np.asarray(self.sizes, dtype=np.int32)
        pointers = np.zeros(len(sizes), dtype=np.int64)
        if len(sizes) > 1:
            np.cumsum(sizes[:-1].astype(np.int64) * self.dtype.itemsize, out=pointers[1:])
        with open(self.prefix + ".idx", "wb") as f:
            f.write(_HDR_MAGIC)
            f.write(struct.pack("<Q", 1))
            f.write(struct.pack("<B", _DTYPE_CODES[self.dtype]))
            f.write(struct.pack("<Q", len(sizes)))
            f.write(struct.pack("<Q", len(self.doc_idx)))
            f.write(sizes.tobytes(order="C"))
            f.write(pointers.tobytes(order="C"))
            f.write(np.asarray(self.doc_idx, dtype=np.int64).tobytes(order="C"))


# =====================================================================================================
# index maps
# =====================================================================================================
def build_sample_idx_numpy(sizes, doc_idx, seq_length, num_epochs, tokens_per_epoch, drop_last=True, add_extra_token=1):
    """Pure-python oracle of the C++ helper."""
    total = num_epochs * tokens_per_epoch - add_extra_token
    num_samples = total // seq_length if drop_last else -(-total // seq_length)
    out = np.zeros((num_samples + 1, 2), dtype=np.int64)
    di, off, s = 0, 0, 1
    while s <= num_samples:
        remaining = seq_length + add_extra_token
        while remaining != 0:
            if di >= len(doc_idx):
                di = len(doc_idx) - 1
                off = int(sizes[doc_idx[di]]) - add_extra_token
                break
            doc_len = int(sizes[doc_idx[di]]) - off
            remaining -= doc_len
            if remaining <= 0:
                off += remaining + doc_len - add_extra_token
                remaining = 0
            else:
                di += 1
                off = 0
        out[s] = (di, off)
        s += 1
    return out


def _num_tokens(documents, sizes) -> int:
    return int(np.sum(sizes[documents]))


def _num_epochs(tokens_per_epoch: int, seq_length: int, num_samples: int, add_extra_token: int = 1) -> int:
    epochs, total = 0, 0
    while True:
        epochs += 1
        total += tokens_per_epoch
        if (total - add_extra_token) // seq_length >= num_samples:
            return epochs


def _build_doc_idx(documents, num_epochs, rng, separate_last_epoch, shuffle=True):
    if not separate_last_epoch or num_epochs == 1:
        doc_idx = np.mgrid[0:num_epochs, 0:len(documents)][1]
        doc_idx[:] = documents
        doc_idx = doc_idx.reshape(-1).astype(np.int64)
        if shuffle:
            rng.shuffle(doc_idx)
        return doc_idx
    first = _build_doc_idx(documents, num_epochs - 1, rng, False, shuffle)
    last = _build_doc_idx(documents, 1, rng, False, shuffle)
    return np.concatenate((first, last))


def _build_shuffle_idx(num_samples: int, total_size: int, rng) -> np.ndarray:
    dtype = np.uint32 if total_size < (np.iinfo(np.uint32).max - 1) else np.int64
    first = np.arange(0, num_samples, dtype=dtype)
    rng.shuffle(first)
    if num_samples == total_size:
        return first
    last = np.arange(num_samples, total_size, dtype=dtype)
    rng.shuffle(last)
    return np.concatenate((first, last))


def build_index_mappings(name: str, data_prefix: str, documents: np.ndarray, sizes: np.ndarray, num_samples: int,
                         seq_length: int, seed: int, index_mapping_dir: Optional[str] = None, drop_last: bool = True,
                         add_extra_token: int = 1, shuffle_documents: bool = True):
    """Returns (doc_idx, sample_idx, shuffle_idx) - built on rank 0, cached as .npy, mmapped by all."""
    tokens_per_epoch = _num_tokens(documents, sizes)
    num_epochs = _num_epochs(tokens_per_epoch, seq_length, num_samples, add_extra_token)
    rng = np.random.RandomState(seed=seed)
    desc = f"{name}_{num_samples}ns_{seq_length}sl_{seed}s_{len(documents)}d_{int(documents[0])}_{int(documents[-1])}"
    tag = hashlib.md5(desc.encode()).hexdigest()[:10]
    base = os.path.join(index_mapping_dir, os.path.basename(data_prefix)) if index_mapping_dir else data_prefix
    base = f"{base}_{name}_indexmap_{tag}"
    paths = [base + s for s in ("_doc_idx.npy", "_sample_idx.npy", "_shuffle_idx.npy")]
    rank0 = (not torch.distributed.is_initialized()) or torch.distributed.get_rank() == 0
    if rank0 and not all(os.path.isfile(p) for p in paths):
        t0 = time.time()
        if index_mapping_dir:
            os.makedirs(index_mapping_dir, exist_ok=True)
        if num_epochs == 1:
            separate_last = False
        else:
            per_epoch = (tokens_per_epoch - add_extra_token) // seq_length
            from_prev = ((num_epochs - 1) * tokens_per_epoch - add_extra_token) // seq_length
            last_epoch_samples = num_samples - from_prev
            separate_last = last_epoch_samples < int(0.80 * per_epoch)
        doc_idx = _build_doc_idx(documents, num_epochs, rng, separate_last, shuffle_documents)
        helpers = helpers_build.get()
        if helpers is not None:
            sample_idx = helpers.build_sample_idx(sizes.astype(np.int32), doc_idx, seq_length, num_epochs, tokens_per_epoch,
                                                  drop_last, add_extra_token)
        else:
            sample_idx = build_sample_idx_numpy(sizes, doc_idx, seq_length, num_epochs, tokens_per_epoch, drop_last, add_extra_token)
        n_shuffle = from_prev if (num_epochs > 1 and separate_last) else sample_idx.shape[0] - 1
        shuffle_idx = _build_shuffle_idx(n_shuffle, sample_idx.shape[0] - 1, rng)
        for p, a in zip(paths, (doc_idx, sample_idx, shuffle_idx)):
            np.save(p, a, allow_pickle=True)
        log.info("built index maps for %s in %.2fs (%d samples, %d epochs)", name, time.time() - t0,
                 sample_idx.shape[0] - 1, num_epochs)
    ps.barrier()
    doc_idx = np.load(paths[0], allow_pickle=True, mmap_mode="r")
    sample_idx = np.load(paths[1], allow_pickle=True, mmap_mode="r")
    shuffle_idx = np.load(paths[2], allow_pickle=True, mmap_mode="r")
    return doc_idx, sample_idx, shuffle_idx


def create_ltor_masks_and_position_ids(tokens: torch.Tensor, eod_token: int, reset_position_ids: bool,
                                       reset_attention_mask: bool, eod_mask_loss: bool):
    """Per-sample loss mask / position ids (and the document-boundary info for attention-mask reset)."""
    S = tokens.numel()
    loss_mask = torch.ones(S, dtype=torch.float)
    if eod_mask_loss:
        loss_mask[tokens == eod_token] = 0.0
    position_ids = torch.arange(S, dtype=torch.long)
    doc_ids = torch.zeros(S, dtype=torch.long)
    if reset_position_ids or reset_attention_mask:
        eod = (tokens == eod_token).nonzero().flatten().tolist()
        prev = 0
        for j, i in enumerate(eod):
            if reset_position_ids:
                position_ids[i + 1:] -= (i + 1 - prev)
                prev = i + 1
            if reset_attention_mask:
                doc_ids[i + 1:] = j + 1
    return loss_mask, position_ids, doc_ids


class GPTDataset(Dataset):
    def __init__(self, name: str, data_prefix: str, documents: np.ndarray, indexed: MMapIndexedDataset, num_samples: int,
                 seq_length: int, seed: int, eod_token: int = 0, reset_position_ids: bool = False,
                 reset_attention_mask: bool = False, eod_mask_loss: bool = False, index_mapping_dir: Optional[str] = None,
                 drop_last: bool = True, no_seqlen_plus_one_input_tokens: bool = False, shuffle_documents: bool = True):
        self.name, self.indexed, self.seq_length = name, indexed, seq_length
        self.eod, self.rpi, self.ram, self.eml = eod_token, reset_position_ids, reset_attention_mask, eod_mask_loss
        self.add_extra_token = 0 if no_seqlen_plus_one_input_tokens else 1
        self.doc_idx, self.sample_idx, self.shuffle_idx = build_index_mappings(
            name, data_prefix, documents, indexed.sizes, num_samples, seq_length, seed, index_mapping_dir, drop_last,
            self.add_extra_token, shuffle_documents)

    def __len__(self):
        return self.sample_idx.shape[0] - 1

    def _tokens(self, idx: int) -> np.ndarray:
        idx = int(self.shuffle_idx[idx])
        d0, o0 = int(self.sample_idx[idx][0]), int(self.sample_idx[idx][1])
        d1, o1 = int(self.sample_idx[idx + 1][0]), int(self.sample_idx[idx + 1][1])
        if d0 == d1:
            return self.indexed.get(int(self.doc_idx[d0]), offset=o0, length=o1 - o0 + self.add_extra_token)
        parts = [self.indexed.get(int(self.doc_idx[d0]), offset=o0)]
        for d in range(d0 + 1, d1):
            parts.append(self.indexed.get(int(self.doc_idx[d])))
        parts.append(self.indexed.get(int(self.doc_idx[d1]), length=o1 + self.add_extra_token))
        return np.concatenate(parts)

    def __getitem__(self, idx: int) -> Dict[str, torch.Tensor]:
        if idx < 0:
            # padding sample (data.pad_samples_to_global_batch_size): a real sample with an all-zero loss mask, so the
            # last partial validation batch has the global batch shape without contributing to the loss
            out = self[0]
            out["loss_mask"] = torch.zeros_like(out["loss_mask"])
            return out
        text = torch.from_numpy(self._tokens(idx).astype(np.int64))
        need = self.seq_length + self.add_extra_token
        if text.numel() < need:                  # last partial sample when drop_last=False
            text = torch.cat([text, torch.full((need - text.numel(),), self.eod, dtype=torch.long)])
        if self.add_extra_token:
            tokens, labels = text[:-1].contiguous(), text[1:].contiguous()
        else:
            tokens = text
            labels = torch.roll(text, shifts=-1, dims=0)
            labels[-1] = -1
        loss_mask, position_ids, doc_ids = create_ltor_masks_and_position_ids(tokens, self.eod, self.rpi, self.ram, self.eml)
        loss_mask[labels == -1] = 0.0
        labels = labels.clamp(min=0)
        out = {"tokens": tokens, "labels": labels, "loss_mask": loss_mask, "position_ids": position_ids}
        if self.ram:
            out["doc_ids"] = doc_ids
        return out


class BlendableDataset(Dataset):
    def __init__(self, datasets: Sequence[Dataset], weights: Sequence[float], size: int):
        self.datasets = list(datasets)
        w = np.asarray(weights, dtype=np.float64)
        w = w / w.sum()
        self.size = size
        self.dataset_index = np.zeros(size, dtype=np.uint8)
        self.dataset_sample_index = np.zeros(size, dtype=np.int64)
        helpers = helpers_build.get()
        if helpers is not None:
            helpers.build_blending_indices(self.dataset_index, self.dataset_sample_index, w, len(datasets), size, False)
        else:
            cur = np.zeros(len(datasets), dtype=np.int64)
            for i in range(size):
                err = w * max(i, 1) - cur
                d = int(np.argmax(err))
                self.dataset_index[i], self.dataset_sample_index[i] = d, cur[d]
                cur[d] += 1

    def __len__(self):
        return self.size

    def __getitem__(self, idx):
        if idx < 0:
            out = self[0]
            out["loss_mask"] = torch.zeros_like(out["loss_mask"])
            return out
        d = int(self.dataset_index[idx])
        return self.datasets[d][int(self.dataset_sample_index[idx]) % len(self.datasets[d])]


def get_train_valid_test_split(splits_string: str, size: int) -> List[int]:
    splits = [float(s) for s in str(splits_string).replace("/", ",").split(",")]
    while len(splits) < 3:
        splits.append(0.0)
    splits = splits[:3]
    total = sum(splits)
    assert total > 0
    splits = [s / total for s in splits]
    idx = [0]
    for s in splits:
        idx.append(idx[-1] + int(round(s * float(size))))
    diff = idx[-1] - size
    for i in range(1, len(idx)):
        idx[i] -= diff
    return idx


def parse_data_prefix(data_prefix) -> Tuple[List[str], List[float]]:
    """``[w1, p1, w2, p2, ...]`` or ``[p]`` or a plain string."""
    if isinstance(data_prefix, str):
        return [data_prefix], [1.0]
    items = list(data_prefix)
    if len(items) == 1:
        return [str(items[0])], [1.0]
    assert len(items) % 2 == 0, "data_prefix must be [weight, prefix, weight, prefix, ...]"
    ws = [float(items[i]) for i in range(0, len(items), 2)]
    ps_ = [str(items[i]) for i in range(1, len(items), 2)]
    return ps_, ws


def build_train_valid_test_datasets(data_prefix, splits_string: str, num_samples: Sequence[int], seq_length: int, seed: int,
                                    **kw) -> Tuple[Optional[Dataset], Optional[Dataset], Optional[Dataset]]:
    prefixes, weights = parse_data_prefix(data_prefix)

    def one(prefix: str, samples: Sequence[int]):
        indexed = MMapIndexedDataset(prefix)
        total_docs = len(indexed.doc_idx) - 1
        splits = get_train_valid_test_split(splits_string, total_docs)
        out = []
        for i, name in enumerate(("train", "valid", "test")):
            if splits[i + 1] > splits[i] and samples[i] > 0:
                docs = np.arange(splits[i], splits[i + 1], dtype=np.int32)
                out.append(GPTDataset(name, prefix, docs, indexed, samples[i], seq_length, seed, **kw))
            else:
                out.append(None)
        return out
    if len(prefixes) == 1:
        return tuple(one(prefixes[0], num_samples))
    wsum = sum(weights)
    per = []
    for p, w in zip(prefixes, weights):
        # slight oversampling so the blend never runs dry
        per.append(one(p, [int(np.ceil(n * (w / wsum) * 1.005)) for n in num_samples]))
    res = []
    for i in range(3):
        dsets = [d[i] for d in per]
        res.append(BlendableDataset(dsets, weights, num_samples[i]) if all(d is not None for d in dsets) and num_samples[i] > 0 else None)
    return tuple(res)


# =====================================================================================================
# tokenizers
# =====================================================================================================
def get_tokenizer(cfg: Optional[Config]):
    """``data.tokenizer.{library,type,model,vocab_file,merge_file}`` (reference megatron/data_module.py:318-339)."""
    if cfg is None:
        return None
    lib = str(cfg.get("library", "huggingface"))
    if lib == "megatron":
        from transformers import GPT2TokenizerFast
        return GPT2TokenizerFast(vocab_file=cfg.get("vocab_file"), merges_file=cfg.get("merge_file"))
    if lib == "sentencepiece":
        import sentencepiece as spm
        sp = spm.SentencePieceProcessor()
        sp.Load(cfg.get("model"))
        sp.legacy = bool(cfg.get("sentencepiece_legacy", False))      # NeMo's flag; special tokens are plain pieces here
        return sp
    from transformers import AutoTokenizer
    return AutoTokenizer.from_pretrained(cfg.get("type", cfg.get("model", None)), use_fast=bool(cfg.get("use_fast", True)))


# =====================================================================================================
# data module
# =====================================================================================================
class MegatronRandomBatchSampler(ResumableBatchSampler):
    def __init__(self, *a, **kw):
        kw["shuffle"] = True
        super().__init__(*a, **kw)


class MegatronDataModule(BaseDataModule):
    def __init__(self, cfg: Config, trainer=None):
        super().__init__(cfg, trainer)
        d = cfg.data
        if d.get("fine_tuning", False):
            raise ValueError("MegatronDataModule is for pre-training; use model_alignment_strategy for fine-tuning")
        self.seq_length = int(d.get("seq_length", cfg.model.get("encoder_seq_length", 2048)))
        self.seed = int(cfg.get("seed", 1234))
        self._ds = None

    def _sample_counts(self) -> List[int]:
        tr = self.cfg.trainer
        max_steps = int(tr.get("max_steps", 0))
        train = max_steps * self.global_batch_size
        vci = tr.get("val_check_interval", None)
        lvb = tr.get("limit_val_batches", 0)
        n_val_runs = (max_steps // int(vci) + 1) if vci and isinstance(vci, int) and vci <= max_steps else 1
        val_batches = int(lvb) if isinstance(lvb, int) or (isinstance(lvb, float) and lvb > 1) else 0
        ltb = tr.get("limit_test_batches", 0)
        test_batches = int(ltb) if isinstance(ltb, int) or (isinstance(ltb, float) and ltb > 1) else 0
        return [train, n_val_runs * val_batches * self.global_batch_size, test_batches * self.global_batch_size]

    def build(self):
        if self._ds is not None:
            return self._ds
        d = self.cfg.data
        tok = d.get("tokenizer", None)
        eod = int(tok.get("eos_id", 0)) if tok is not None and tok.get("eos_id", None) is not None else 0
        self._ds = build_train_valid_test_datasets(
            d.data_prefix, str(d.get("splits_string", "980,10,10")), self._sample_counts(), self.seq_length, self.seed,
            eod_token=eod, reset_position_ids=bool(d.get("reset_position_ids", False)),
            reset_attention_mask=bool(d.get("reset_attention_mask", False)), eod_mask_loss=bool(d.get("eod_mask_loss", False)),
            index_mapping_dir=d.get("index_mapping_dir", None), drop_last=True,
            no_seqlen_plus_one_input_tokens=bool(d.get("no_seqlen_plus_one_input_tokens", False)))
        return self._ds

    def _loader(self, ds, consumed: int, drop_last: bool = True, pad: bool = False):
        if ds is None:
            return None
        d = self.cfg.data
        impl = str(d.get("data_impl", "mmap"))
        if impl not in ("mmap", "lazy", "cached", "infer"):
            raise ValueError(f"data.data_impl={impl!r}: only Megatron indexed datasets (.bin/.idx, memory-mapped) are supported")
        # data.skip_warmup: the memory-mapped reader never pre-touches the file, so the flag has nothing to skip
        cls = MegatronRandomBatchSampler if str(d.get("dataloader_type", "single")) == "cyclic" else ResumableBatchSampler
        sampler = cls(len(ds), consumed, self.global_batch_size, self.dp_rank, self.dp_size, drop_last=drop_last, seed=self.seed)
        if pad and not drop_last and hasattr(sampler, "pad_to_global_batch"):
            sampler.pad_to_global_batch = True
        return DataLoader(ds, batch_sampler=sampler, collate_fn=default_collate, num_workers=int(d.get("num_workers", 0)),
                          pin_memory=torch.cuda.is_available(), persistent_workers=int(d.get("num_workers", 0)) > 0)

    def train_dataloader(self):
        return self._loader(self.build()[0], self.consumed_samples)

    def val_dataloader(self):
        d = self.cfg.data
        # reference megatron/data_module.py:62-75: the last partial validation batch is dropped unless
        # validation_drop_last is False, in which case it may be padded up to the global batch
        return self._loader(self.build()[1], 0, drop_last=bool(d.get("validation_drop_last", True)),
                            pad=bool(d.get("pad_samples_to_global_batch_size", False)))

    def test_dataloader(self):
        return self._loader(self.build()[2], 0)

    def get_batch_length(self, batch):
        return len(batch["tokens"])
