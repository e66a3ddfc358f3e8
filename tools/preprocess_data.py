#!/usr/bin/env python
"""Tokenise raw text for pre-training.

Two output formats, matching the two pre-training data modules:

  megatron   ``<prefix>.bin`` / ``<prefix>.idx`` memory-mapped indexed dataset (``data.data_prefix``; the reference relies
             on NeMo's ``preprocess_data_for_megatron.py`` for this): one document per input record, optional EOD append
  hf         a ``datasets`` arrow directory of fixed-length ``input_ids`` / ``attention_mask`` / ``labels`` blocks
             (``data.train_dir`` of the HF data module; documents concatenated with EOS and cut into ``--seq-length``)

    python tools/preprocess_data.py --input corpus.jsonl --json-key text --format megatron --output-prefix data/corpus \\
        --tokenizer-library huggingface --tokenizer-type /models/llama3_tokenizer --append-eod
    python tools/preprocess_data.py --input corpus.jsonl --format hf --output-dir data/corpus_arrow --seq-length 8192 \\
        --tokenizer-type /models/llama3_tokenizer
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from typing import Iterable, Iterator, List

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def read_records(path: str, key: str) -> Iterator[str]:
    with open(path) as f:
        if path.endswith(".jsonl") or path.endswith(".json"):
            for line in f:
                line = line.strip()
                if line:
                    yield str(json.loads(line)[key])
        else:                                   # plain text: one document per line
            for line in f:
                if line.strip():
                    yield line.rstrip("\n")


def load_tokenizer(args):
    from neuronx_distributed_training_b200.config import from_dict
    from neuronx_distributed_training_b200.data.megatron import get_tokenizer
    return get_tokenizer(from_dict({"library": args.tokenizer_library, "type": args.tokenizer_type, "model": args.tokenizer_model,
                                    "vocab_file": args.vocab_file, "merge_file": args.merge_file}))


def encode(tok, text: str) -> List[int]:
    if hasattr(tok, "EncodeAsIds"):             # sentencepiece
        return list(tok.EncodeAsIds(text))
    try:
        return list(tok.encode(text, add_special_tokens=False))
    except TypeError:
        return list(tok.encode(text))


def eod_id(tok, override) -> int:
    if override is not None:
        return int(override)
    for name in ("eos_token_id", "eod", "eos_id"):
        v = getattr(tok, name, None)
        if callable(v):
            v = v()
        if v is not None:
            return int(v)
    return 0


def write_megatron(docs: Iterable[List[int]], prefix: str, append_eod: bool, eod: int, vocab_size: int) -> dict:
    import numpy as np
    from neuronx_distributed_training_b200.data.megatron import MMapIndexedDatasetBuilder
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    b = MMapIndexedDatasetBuilder(prefix, dtype=np.uint16 if vocab_size and vocab_size < 65500 else np.int32)
    n_docs = n_tok = 0
    for ids in docs:
        if append_eod:
            ids = ids + [eod]
        if not ids:
            continue
        b.add_item(ids)
        b.end_document()
        n_docs += 1
        n_tok += len(ids)
    b.finalize()
    return {"documents": n_docs, "tokens": n_tok, "prefix": prefix}


def write_hf(docs: Iterable[List[int]], out_dir: str, seq_length: int, eos: int) -> dict:
    import datasets
    buf: List[int] = []
    rows = []
    n_docs = 0
    for ids in docs:
        n_docs += 1
        buf.extend(ids + [eos])
        while len(buf) >= seq_length:
            chunk, buf = buf[:seq_length], buf[seq_length:]
            rows.append({"input_ids": chunk, "attention_mask": [1] * seq_length, "labels": list(chunk)})
    ds = datasets.Dataset.from_list(rows)
    ds.save_to_disk(out_dir)
    return {"documents": n_docs, "sequences": len(rows), "dropped_tail_tokens": len(buf), "dir": out_dir}


def main(argv=None) -> dict:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--input", required=True)
    ap.add_argument("--json-key", default="text")
    ap.add_argument("--format", choices=["megatron", "hf"], default="megatron")
    ap.add_argument("--output-prefix")
    ap.add_argument("--output-dir")
    ap.add_argument("--seq-length", type=int, default=8192)
    ap.add_argument("--append-eod", action="store_true")
    ap.add_argument("--eod-id", type=int, default=None)
    ap.add_argument("--tokenizer-library", default="huggingface", choices=["huggingface", "megatron", "sentencepiece"])
    ap.add_argument("--tokenizer-type", default=None, help="HF name / path (huggingface) ")
    ap.add_argument("--tokenizer-model", default=None, help="sentencepiece model file")
    ap.add_argument("--vocab-file", default=None)
    ap.add_argument("--merge-file", default=None)
    args = ap.parse_args(argv)
    tok = getattr(args, "_tokenizer", None) or load_tokenizer(args)
    docs = (encode(tok, t) for t in read_records(args.input, args.json_key))
    eod = eod_id(tok, args.eod_id)
    if args.format == "megatron":
        assert args.output_prefix, "--output-prefix is required for the megatron format"
        vs = getattr(tok, "vocab_size", None)
        vs = vs() if callable(vs) else vs
        info = write_megatron(docs, args.output_prefix, args.append_eod, eod, int(vs or 0))
    else:
        assert args.output_dir, "--output-dir is required for the hf format"
        info = write_hf(docs, args.output_dir, args.seq_length, eod)
    print(json.dumps(info))
    return info


if __name__ == "__main__":
    main()
