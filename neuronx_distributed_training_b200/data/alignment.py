"""Model-alignment data: SFT (packing or padding), DPO and ORPO preference pairs.

Capability parity: reference model_alignment_data_module.py:34-255, datasets/ConcatDataset.py:7-81,
datasets/PaddedDataset.py:9-103, data/utils.py:6-64.  The ``trl`` preference tokenisation the reference borrows is
re-implemented here (prompt / chosen / rejected tokenised separately, prompt truncated per ``truncation_mode`` to
``max_prompt_length``, answers truncated to fit ``seq_length``, labels = -100 on the prompt).
Packing fixes the reference's bug (every key was extended with ``input_ids``): each key is packed
from its own column, so the prompt mask survives packing.
"""
from __future__ import annotations

import json
import logging
import os
from typing import Any, Dict, List, Optional, Sequence

import torch
from torch.utils.data import DataLoader, Dataset, DistributedSampler

from ..config import Config
from .base import BaseDataModule, default_collate

log = logging.getLogger("nxdt")
IGNORE_INDEX = -100


# ---- loading -------------------------------------------------------------------------------------------
def load_records(path: str, dev_choose_samples: Optional[int] = None) -> List[Dict[str, Any]]:
    """``.jsonl`` / ``.json`` / HF ``datasets`` arrow directory."""
    if os.path.isdir(path) or path.endswith(".arrow"):
        import datasets
        ds = datasets.load_from_disk(path) if os.path.isdir(path) else datasets.Dataset.from_file(path)
        if hasattr(ds, "keys") and "train" in ds:
            ds = ds["train"]
        recs = [dict(r) for r in ds]
    elif path.endswith(".jsonl"):
        with open(path) as f:
            recs = [json.loads(l) for l in f if l.strip()]
    elif path.endswith(".json"):
        with open(path) as f:
            obj = json.load(f)
        recs = obj if isinstance(obj, list) else obj.get("data", obj.get("train", []))
    else:
        raise ValueError(f"unsupported alignment data file {path!r} (expected .jsonl/.json/.arrow or a datasets dir)")
    if dev_choose_samples:
        recs = recs[: int(dev_choose_samples)]
    return recs


def apply_prompt_template(recs: List[Dict[str, Any]], template: Optional[str]) -> List[Dict[str, Any]]:
    """Minimal jinja-free templating (``{field}`` placeholders) replacing the reference's optional promptsource step."""
    if not template:
        return recs
    out = []
    for r in recs:
        rr = dict(r)
        rr["input"] = template.format(**r)
        out.append(rr)
    return out


# ---- tokenisation --------------------------------------------------------------------------------------
def _ids(tokenizer, text: str) -> List[int]:
    if hasattr(tokenizer, "encode"):
        try:
            return list(tokenizer.encode(text, add_special_tokens=False))
        except TypeError:
            return list(tokenizer.encode(text))
    return list(tokenizer(text, add_special_tokens=False)["input_ids"])


def tokenize_sft(rec: Dict[str, Any], tokenizer, seq_length: int, input_key: str = "input", output_key: str = "output") -> Dict[str, List[int]]:
    """``bos + input || output + eos``; labels are -100 on the prompt part (reference :148-160)."""
    bos, eos = getattr(tokenizer, "bos_token_id", None), getattr(tokenizer, "eos_token_id", None)
    p = ([bos] if bos is not None else []) + _ids(tokenizer, str(rec.get(input_key, rec.get("prompt", ""))))
    a = _ids(tokenizer, str(rec.get(output_key, rec.get("completion", rec.get("response", ""))))) + ([eos] if eos is not None else [])
    ids = (p + a)[:seq_length]
    labels = ([IGNORE_INDEX] * len(p) + a)[:seq_length]
    return {"input_ids": ids, "attention_mask": [1] * len(ids), "labels": labels}


def tokenize_preference(rec: Dict[str, Any], tokenizer, seq_length: int, max_prompt_length: int,
                        truncation_mode: str = "keep_end") -> Dict[str, List[int]]:
    """prompt / chosen / rejected -> chosen_* and rejected_* sequences sharing the (truncated) prompt."""
    bos, eos = getattr(tokenizer, "bos_token_id", None), getattr(tokenizer, "eos_token_id", None)
    prompt = _ids(tokenizer, str(rec["prompt"]))
    chosen = _ids(tokenizer, str(rec["chosen"]))
    rejected = _ids(tokenizer, str(rec["rejected"]))
    if bos is not None:
        prompt = [bos] + prompt
    if eos is not None:
        chosen, rejected = chosen + [eos], rejected + [eos]
    longer = max(len(chosen), len(rejected))
    if len(prompt) + longer > seq_length:
        if truncation_mode == "keep_start":
            prompt = prompt[:max_prompt_length]
        elif truncation_mode == "keep_end":
            prompt = prompt[-max_prompt_length:]
        else:
            raise ValueError(f"unknown truncation_mode {truncation_mode}")
    if len(prompt) + longer > seq_length:
        room = seq_length - len(prompt)
        chosen, rejected = chosen[:room], rejected[:room]
    out = {"prompt_input_ids": prompt, "prompt_attention_mask": [1] * len(prompt)}
    for name, ans in (("chosen", chosen), ("rejected", rejected)):
        out[f"{name}_input_ids"] = prompt + ans
        out[f"{name}_attention_mask"] = [1] * (len(prompt) + len(ans))
        out[f"{name}_labels"] = [IGNORE_INDEX] * len(prompt) + ans
    return out


# ---- datasets --------------------------------------------------------------------------------------------
class PaddedDataset(Dataset):
    """right-pad every SFT sample to ``seq_length``."""

    def __init__(self, samples: List[Dict[str, List[int]]], seq_length: int, pad_id: int):
        self.samples, self.seq_length, self.pad_id = samples, seq_length, pad_id

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        s = self.samples[i]
        n = self.seq_length - len(s["input_ids"])
        return {"input_ids": torch.tensor(s["input_ids"] + [self.pad_id] * n), "attention_mask": torch.tensor(s["attention_mask"] + [0] * n),
                "labels": torch.tensor(s["labels"] + [IGNORE_INDEX] * n)}


class ConcatDataset(Dataset):
    """Greedy packing: samples are concatenated (each key from its own column) into ``seq_length`` rows; a sample
    that does not fit starts a new row, leftovers are padded."""

    def __init__(self, samples: List[Dict[str, List[int]]], seq_length: int, pad_id: int):
        self.rows: List[Dict[str, List[int]]] = []
        cur = {"input_ids": [], "attention_mask": [], "labels": []}
        pads = {"input_ids": pad_id, "attention_mask": 0, "labels": IGNORE_INDEX}

        def flush():
            if cur["input_ids"]:
                n = seq_length - len(cur["input_ids"])
                self.rows.append({k: v + [pads[k]] * n for k, v in cur.items()})
                for v in cur.values():
                    v.clear()
        for s in samples:
            L = len(s["input_ids"])
            if L > seq_length:
                s = {k: v[:seq_length] for k, v in s.items()}
                L = seq_length
            if len(cur["input_ids"]) + L > seq_length:
                flush()
            for k in cur:
                cur[k].extend(s[k])
            # never learn to predict the first token of the next packed sample from the previous one
            cur["labels"][-L] = IGNORE_INDEX
        flush()

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, i):
        return {k: torch.tensor(v) for k, v in self.rows[i].items()}


class PaddedDPODataset(Dataset):
    """chosen / rejected padded to ``seq_length`` (right), prompt left-padded to ``max_prompt_length``
    (reference PaddedDataset.py:42-97); carries reference log-probs once they are pre-computed."""

    def __init__(self, samples: List[Dict[str, List[int]]], seq_length: int, max_prompt_length: int, pad_id: int):
        self.samples, self.seq_length, self.max_prompt_length, self.pad_id = samples, seq_length, max_prompt_length, pad_id
        self.extra: Dict[str, List[float]] = {}

    def add_column(self, name: str, values: Sequence[float]):
        assert len(values) == len(self.samples)
        self.extra[name] = list(values)

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        s = self.samples[i]
        out = {}
        for name in ("chosen", "rejected"):
            n = self.seq_length - len(s[f"{name}_input_ids"])
            out[f"{name}_input_ids"] = torch.tensor(s[f"{name}_input_ids"] + [self.pad_id] * n)
            out[f"{name}_attention_mask"] = torch.tensor(s[f"{name}_attention_mask"] + [0] * n)
            out[f"{name}_labels"] = torch.tensor(s[f"{name}_labels"] + [IGNORE_INDEX] * n)
        p = s["prompt_input_ids"][-self.max_prompt_length:]
        n = self.max_prompt_length - len(p)
        out["prompt_input_ids"] = torch.tensor([self.pad_id] * n + p)
        out["prompt_attention_mask"] = torch.tensor([0] * n + [1] * len(p))
        for k, v in self.extra.items():
            out[k] = torch.tensor(v[i], dtype=torch.float32)
        return out


# ---- data module --------------------------------------------------------------------------------------------
class ModelAlignmentDataModule(BaseDataModule):
    def __init__(self, cfg: Config, trainer=None, tokenizer=None):
        super().__init__(cfg, trainer)
        a = cfg.model_alignment_strategy
        self.mode = "dpo" if a.get("dpo", None) is not None else "orpo" if a.get("orpo", None) is not None else "sft"
        self.strategy_cfg = a.get(self.mode, Config({})) or Config({})
        self.seq_length = int(cfg.data.get("seq_length", cfg.model.get("encoder_seq_length", 4096)))
        self.packing = bool((a.get("sft", Config({})) or Config({})).get("packing", False)) if self.mode == "sft" else False
        self.max_prompt_length = int(self.strategy_cfg.get("max_prompt_length", self.seq_length // 2))
        self.truncation_mode = str(self.strategy_cfg.get("truncation_mode", "keep_start"))
        self.seed = int(cfg.get("seed", 1234))
        self._tokenizer = tokenizer
        self._train = self._val = None

    @property
    def tokenizer(self):
        if self._tokenizer is None:
            from transformers import AutoTokenizer
            t = self.cfg.data.tokenizer
            self._tokenizer = AutoTokenizer.from_pretrained(t.get("type", t.get("model", None)))
            if self._tokenizer.pad_token_id is None:
                self._tokenizer.pad_token = self._tokenizer.eos_token
        return self._tokenizer

    def _pad_id(self) -> int:
        t = self.tokenizer
        pid = getattr(t, "pad_token_id", None)
        return int(pid if pid is not None else getattr(t, "eos_token_id", 0) or 0)

    def _build(self, path: Optional[str]):
        if not path:
            return None
        recs = load_records(path, self.cfg.data.get("dev_choose_samples", None))
        recs = apply_prompt_template(recs, self.cfg.data.get("prompt_template", None))
        if self.mode == "sft":
            samples = [tokenize_sft(r, self.tokenizer, self.seq_length) for r in recs]
            return ConcatDataset(samples, self.seq_length, self._pad_id()) if self.packing \
                else PaddedDataset(samples, self.seq_length, self._pad_id())
        samples = [tokenize_preference(r, self.tokenizer, self.seq_length, self.max_prompt_length, self.truncation_mode) for r in recs]
        return PaddedDPODataset(samples, self.seq_length, self.max_prompt_length, self._pad_id())

    def setup(self, ckpt_path=None):
        super().setup(ckpt_path)
        d = self.cfg.data
        self._train = self._build(d.get("train_dir", None))
        self._val = self._build(d.get("val_dir", None))

    def train_dataset(self):
        return self._train

    def _loader(self, ds, shuffle: bool, epoch: int = 0):
        if ds is None:
            return None
        sampler = DistributedSampler(ds, num_replicas=self.dp_size, rank=self.dp_rank, shuffle=shuffle, seed=self.seed, drop_last=True)
        sampler.set_epoch(epoch)
        return DataLoader(ds, batch_size=self.per_dp_batch, sampler=sampler, collate_fn=default_collate, num_workers=0,
                          pin_memory=torch.cuda.is_available(), drop_last=True)

    def train_dataloader(self, shuffle: bool = True):
        epoch = self.trainer.current_epoch if self.trainer is not None else 0
        return self._loader(self._train, shuffle, epoch)

    def val_dataloader(self):
        return self._loader(self._val, False)

    def get_batch_length(self, batch):
        key = "input_ids" if "input_ids" in batch else "chosen_input_ids"
        return len(batch[key])
