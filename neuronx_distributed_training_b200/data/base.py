"""Data module base + synthetic and HF-arrow pre-training modules.

Capability parity: reference lightning_modules/data/base.py:14-95 (DP size, num_microbatches,
``consumed_samples`` parsed from the checkpoint name and advanced per step, batches kept on the host)
and hf_data_module.py:14-59 (pre-tokenised arrow dataset, per-DP-rank sampler, default collate).
Batches stay on (pinned) host memory; the trainer issues the H2D copies on a side stream.
"""
from __future__ import annotations

from typing import Any, Dict, Iterator, List, Optional

import torch
from torch.utils.data import DataLoader, Dataset, Sampler

from ..config import Config
from ..parallel import state as ps
from ..utils.exp_manager import consumed_samples_from_path


class BaseDataModule:
    def __init__(self, cfg: Config, trainer=None):
        self.cfg = cfg
        self.config = cfg.data
        self.trainer = trainer
        self.dp_size = ps.get_data_parallel_size()
        self.dp_rank = ps.get_data_parallel_rank()
        self.micro_batch_size = int(self.config.micro_batch_size)
        self.global_batch_size = int(self.config.global_batch_size)
        assert self.global_batch_size % (self.micro_batch_size * self.dp_size) == 0, \
            "global_batch_size must be divisible by micro_batch_size * data_parallel_size"
        self.num_microbatches = self.global_batch_size // (self.micro_batch_size * self.dp_size)
        self.per_dp_batch = self.global_batch_size // self.dp_size
        self.init_consumed_samples = 0
        self.consumed_samples = 0

    # ---- lifecycle ------------------------------------------------------------------------------
    def setup(self, ckpt_path: Optional[str] = None):
        self.init_consumed_samples = consumed_samples_from_path(ckpt_path)
        self.consumed_samples = self.init_consumed_samples

    def compute_consumed_samples(self, steps_since_resume: int) -> int:
        return self.init_consumed_samples + steps_since_resume * self.global_batch_size

    def get_batch_length(self, batch: Dict[str, torch.Tensor]) -> int:
        return len(batch["input_ids"])

    def process_global_batch(self, global_batch, global_batch_size=None):
        return global_batch

    def train_dataloader(self):
        raise NotImplementedError

    def val_dataloader(self):
        return None

    def test_dataloader(self):
        return None

    def state_dict(self) -> Dict[str, Any]:
        return {"consumed_samples": self.consumed_samples}

    def load_state_dict(self, sd: Dict[str, Any]):
        self.consumed_samples = int(sd.get("consumed_samples", 0))
        self.init_consumed_samples = self.consumed_samples


def pin_batch(batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    if not torch.cuda.is_available():
        return batch
    return {k: (v.pin_memory() if torch.is_tensor(v) and not v.is_pinned() else v) for k, v in batch.items()}


# ---------------------------------------------------------------------------------------------------
class SyntheticTokenDataset(Dataset):
    """Deterministic random tokens: sample i is a function of (seed, i) only, so every parallel layout
    and every resume sees the same stream."""

    def __init__(self, vocab_size: int, seq_length: int, num_samples: int, seed: int = 1234):
        self.vocab_size, self.seq_length, self.num_samples, self.seed = vocab_size, seq_length, num_samples, seed

    def __len__(self):
        return self.num_samples

    def __getitem__(self, i: int):
        g = torch.Generator()
        g.manual_seed(self.seed * 1000003 + int(i))
        ids = torch.randint(0, self.vocab_size, (self.seq_length,), generator=g, dtype=torch.long)
        return {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": ids.clone()}


class ResumableBatchSampler(Sampler):
    """Yields this DP rank's slice of each global batch, starting after ``consumed_samples``.

    Same contract as the Megatron pre-training batch sampler the reference uses
    (reference megatron/data_module.py:132-173): global sample order is 0,1,2,...; global batch ``b`` is
    samples ``[b*GBS, (b+1)*GBS)``; DP rank ``r`` takes the contiguous ``[r*GBS/DP, (r+1)*GBS/DP)`` slice.
    """

    def __init__(self, total_samples: int, consumed_samples: int, global_batch_size: int, dp_rank: int, dp_size: int,
                 drop_last: bool = True, shuffle: bool = False, seed: int = 0):
        self.total, self.consumed = total_samples, consumed_samples
        self.gbs, self.rank, self.size = global_batch_size, dp_rank, dp_size
        self.per_rank = global_batch_size // dp_size
        self.drop_last, self.shuffle, self.seed = drop_last, shuffle, seed
        self.pad_to_global_batch = False      # with drop_last=False: fill the last batch with index -1 (padding samples)

    @property
    def epoch_span(self) -> int:
        """Samples the ``consumed_samples`` counter advances by per epoch: the counter moves one global batch per
        step, so an epoch spans whole global batches (the remainder is dropped, or padded when ``drop_last=False``)."""
        full = (self.total // self.gbs) * self.gbs
        if self.drop_last or full == self.total:
            return max(full, 1)
        return full + self.gbs

    def __len__(self):
        offset = self.consumed % self.epoch_span
        return (self.epoch_span - offset) // self.gbs

    def _order(self, epoch: int) -> List[int]:
        if not self.shuffle:
            return list(range(self.total))
        g = torch.Generator()
        g.manual_seed(self.seed + epoch)
        return torch.randperm(self.total, generator=g).tolist()

    def __iter__(self) -> Iterator[List[int]]:
        epoch = self.consumed // self.epoch_span
        offset = self.consumed % self.epoch_span
        order = self._order(epoch)
        i = offset
        while i + self.gbs <= self.total:
            lo = i + self.rank * self.per_rank
            yield order[lo: lo + self.per_rank]
            i += self.gbs
        if not self.drop_last and i < self.total:
            rem = order[i:]
            if self.pad_to_global_batch:
                rem = rem + [-1] * (self.gbs - len(rem))
                yield rem[self.rank * self.per_rank: (self.rank + 1) * self.per_rank]
            else:
                per = (len(rem) + self.size - 1) // self.size
                yield rem[self.rank * per: (self.rank + 1) * per]


def default_collate(samples: List[Dict[str, Any]]) -> Dict[str, torch.Tensor]:
    out = {}
    for k in samples[0]:
        v0 = samples[0][k]
        if torch.is_tensor(v0):
            out[k] = torch.stack([s[k] for s in samples])
        elif isinstance(v0, (list, tuple)) or hasattr(v0, "__array__"):
            out[k] = torch.stack([torch.as_tensor(s[k]) for s in samples])
        elif isinstance(v0, (int, float)):
            out[k] = torch.tensor([s[k] for s in samples])
        else:
            out[k] = [s[k] for s in samples]
    return out


class SyntheticDataModule(BaseDataModule):
    """Random-token pre-training data of the configured shape (no files needed)."""

    def __init__(self, cfg: Config, trainer=None, vocab_size: Optional[int] = None, seq_length: Optional[int] = None):
        super().__init__(cfg, trainer)
        self.vocab_size = int(vocab_size or cfg.data.get("vocab_size", None) or cfg.model.get("vocab_size", 32000))
        self.seq_length = int(seq_length or cfg.data.get("seq_length", None) or cfg.model.get("encoder_seq_length"))
        max_steps = int(cfg.trainer.get("max_steps", 1000))
        self.num_samples = int(cfg.data.get("num_samples", max(max_steps, 1) * self.global_batch_size))
        self.seed = int(cfg.get("seed", 1234))

    def _loader(self, seed_offset: int = 0):
        ds = SyntheticTokenDataset(self.vocab_size, self.seq_length, self.num_samples, self.seed + seed_offset)
        sampler = ResumableBatchSampler(len(ds), self.consumed_samples if seed_offset == 0 else 0,
                                        self.global_batch_size, self.dp_rank, self.dp_size)
        return DataLoader(ds, batch_sampler=sampler, collate_fn=default_collate, num_workers=0,
                          pin_memory=torch.cuda.is_available())

    def train_dataloader(self):
        return self._loader(0)

    def val_dataloader(self):
        return self._loader(7)


class HFDataModule(BaseDataModule):
    """Pre-tokenised HF ``datasets`` arrow directory (``datasets.load_from_disk``)."""

    def __init__(self, cfg: Config, trainer=None):
        super().__init__(cfg, trainer)
        self.train_dir = cfg.data.get("train_dir", None)
        self.val_dir = cfg.data.get("val_dir", None)
        self.test_dir = cfg.data.get("test_dir", None)

    def _load(self, path):
        import datasets
        ds = datasets.load_from_disk(path)
        if isinstance(ds, datasets.DatasetDict):
            ds = ds["train"]
        cols = [c for c in ("input_ids", "attention_mask", "labels") if c in ds.column_names]
        ds = ds.with_format("torch", columns=cols)
        return ds

    def _loader(self, path, consumed: int):
        if path is None:
            return None
        ds = self._load(path)
        sampler = ResumableBatchSampler(len(ds), consumed, self.global_batch_size, self.dp_rank, self.dp_size,
                                        drop_last=True, shuffle=False)

        def collate(samples):
            b = default_collate(samples)
            if "labels" not in b:
                b["labels"] = b["input_ids"].clone()
            return b
        return DataLoader(ds, batch_sampler=sampler, collate_fn=collate, num_workers=0,
                          pin_memory=torch.cuda.is_available())

    def train_dataloader(self):
        return self._loader(self.train_dir, self.consumed_samples)

    def val_dataloader(self):
        return self._loader(self.val_dir, 0)

    def test_dataloader(self):
        return self._loader(self.test_dir, 0)


def get_batch_on_this_context_parallel_rank(batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Slice every ``[B, S]`` tensor to this CP rank's contiguous ``S/CP`` chunk.
    Labels are shifted *before* the split so the model must not shift again under CP."""
    cp, r = ps.get_context_model_parallel_size(), ps.get_context_model_parallel_rank()
    if cp == 1:
        return batch
    out = dict(batch)
    if "labels" in out and torch.is_tensor(out["labels"]):
        lab = out["labels"]
        pad = torch.full((lab.shape[0], 1), -100, dtype=lab.dtype, device=lab.device)
        out["labels"] = torch.cat([lab[:, 1:], pad], dim=1)
    if "loss_mask" in out and torch.is_tensor(out["loss_mask"]):
        lm = out["loss_mask"]
        out["loss_mask"] = torch.cat([lm[:, 1:], torch.zeros_like(lm[:, :1])], dim=1)
    for k, v in out.items():
        if torch.is_tensor(v) and v.dim() >= 2:
            S = v.shape[1]
            assert S % cp == 0, f"sequence length {S} not divisible by context parallel size {cp}"
            c = S // cp
            out[k] = v[:, r * c:(r + 1) * c].contiguous()
    return out
