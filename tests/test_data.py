import os
import numpy as np
import pytest
import torch

from neuronx_distributed_training_b200.data import helpers_build
from neuronx_distributed_training_b200.data.megatron import (MMapIndexedDataset, MMapIndexedDatasetBuilder, build_sample_idx_numpy,
                                                             GPTDataset, BlendableDataset, get_train_valid_test_split,
                                                             build_train_valid_test_datasets, create_ltor_masks_and_position_ids)
from neuronx_distributed_training_b200.data.base import ResumableBatchSampler, SyntheticTokenDataset
from neuronx_distributed_training_b200.utils.exp_manager import consumed_samples_from_path, step_from_path


def _make_corpus(tmp_path, n_docs=37, seed=0, eod=0):
    rng = np.random.RandomState(seed)
    prefix = str(tmp_path / "corpus")
    b = MMapIndexedDatasetBuilder(prefix, dtype=np.uint16)
    docs = []
    for _ in range(n_docs):
        n = int(rng.randint(3, 40))
        toks = rng.randint(1, 1000, size=n).tolist() + [eod]
        b.add_item(toks); b.end_document(); docs.append(toks)
    b.finalize()
    return prefix, docs


def test_indexed_dataset_roundtrip(tmp_path):
    prefix, docs = _make_corpus(tmp_path)
    ds = MMapIndexedDataset(prefix)
    assert len(ds) == len(docs) and ds.dtype == np.uint16
    for i in (0, 5, len(docs) - 1):
        assert ds[i].tolist() == docs[i]
    assert ds.get(3, offset=1, length=2).tolist() == docs[3][1:3]


def test_cpp_sample_idx_matches_numpy_oracle(tmp_path):
    helpers = helpers_build.get()
    assert helpers is not None, "C++ data helpers failed to build"
    rng = np.random.RandomState(1)
    sizes = rng.randint(1, 50, size=200).astype(np.int32)
    for epochs in (1, 3):
        doc_idx = np.concatenate([rng.permutation(200) for _ in range(epochs)]).astype(np.int64)
        tpe = int(sizes.sum())
        for seq in (8, 33, 128):
            for drop_last in (True, False):
                a = helpers.build_sample_idx(sizes, doc_idx, seq, epochs, tpe, drop_last, 1)
                b = build_sample_idx_numpy(sizes, doc_idx, seq, epochs, tpe, drop_last, 1)
                assert np.array_equal(a, b)


def test_gpt_dataset_samples_are_contiguous_stream(tmp_path):
    prefix, docs = _make_corpus(tmp_path)
    indexed = MMapIndexedDataset(prefix)
    ds = GPTDataset("train", prefix, np.arange(len(docs), dtype=np.int32), indexed, num_samples=50, seq_length=16, seed=7,
                    shuffle_documents=False)
    assert len(ds) >= 50
    flat = np.concatenate([np.asarray(d) for d in docs] * 4)
    inv = np.argsort(np.asarray(ds.shuffle_idx))        # sample k of the un-shuffled stream
    for k in (0, 1, 7):
        s = ds[int(inv[k])]
        assert s["tokens"].tolist() == flat[k * 16: k * 16 + 16].tolist()
        assert s["labels"].tolist() == flat[k * 16 + 1: k * 16 + 17].tolist()
        assert s["loss_mask"].shape == (16,) and s["position_ids"].tolist() == list(range(16))
    # cache files are reused
    ds2 = GPTDataset("train", prefix, np.arange(len(docs), dtype=np.int32), indexed, 50, 16, 7, shuffle_documents=False)
    assert np.array_equal(np.asarray(ds.shuffle_idx), np.asarray(ds2.shuffle_idx))


def test_eod_resets():
    toks = torch.tensor([5, 6, 0, 7, 8, 9, 0, 3])
    lm, pos, doc = create_ltor_masks_and_position_ids(toks, 0, True, True, True)
    assert lm.tolist() == [1, 1, 0, 1, 1, 1, 0, 1]
    assert pos.tolist() == [0, 1, 2, 0, 1, 2, 3, 0]
    assert doc.tolist() == [0, 0, 0, 1, 1, 1, 1, 2]


def test_splits_and_blending(tmp_path):
    assert get_train_valid_test_split("980,10,10", 1000) == [0, 980, 990, 1000]
    assert get_train_valid_test_split("1", 10) == [0, 10, 10, 10]
    p1, _ = _make_corpus(tmp_path / "a" if (tmp_path / "a").mkdir() is None else tmp_path, 40, 1)
    p2, _ = _make_corpus(tmp_path / "b" if (tmp_path / "b").mkdir() is None else tmp_path, 40, 2)
    tr, va, te = build_train_valid_test_datasets([0.7, p1, 0.3, p2], "90,5,5", [100, 10, 0], 8, 3)
    assert isinstance(tr, BlendableDataset) and len(tr) == 100 and te is None
    frac = (tr.dataset_index == 0).mean()
    assert abs(frac - 0.7) < 0.02
    assert tr[5]["tokens"].shape == (8,)


def test_resumable_sampler_partition_and_resume():
    full = list(ResumableBatchSampler(64, 0, 8, 0, 1))
    r0 = list(ResumableBatchSampler(64, 0, 8, 0, 2)); r1 = list(ResumableBatchSampler(64, 0, 8, 1, 2))
    for b, a0, a1 in zip(full, r0, r1):
        assert b == a0 + a1
    resumed = list(ResumableBatchSampler(64, 24, 8, 1, 2))
    assert resumed == r1[3:]
    ds = SyntheticTokenDataset(100, 16, 10, seed=3)
    assert torch.equal(ds[4]["input_ids"], SyntheticTokenDataset(100, 16, 10, seed=3)[4]["input_ids"])


def test_ckpt_name_parsing():
    p = "/x/checkpoints/hf_llama3_8B--step=10-consumed_samples=10240.0.ckpt"
    assert consumed_samples_from_path(p) == 10240 and step_from_path(p) == 10
    assert consumed_samples_from_path(None) == 0


def test_validation_padding_samples(tmp_path):
    """data.validation_drop_last=False + pad_samples_to_global_batch_size: the last partial batch is filled with index -1,
    which the dataset turns into a sample that carries no loss."""
    from neuronx_distributed_training_b200.data.base import ResumableBatchSampler
    prefix, docs = _make_corpus(tmp_path)
    indexed = MMapIndexedDataset(prefix)
    ds = GPTDataset("valid", prefix, np.arange(len(docs), dtype=np.int32), indexed, num_samples=10, seq_length=16, seed=7)
    pad = ds[-1]
    assert pad["tokens"].shape == ds[0]["tokens"].shape and float(pad["loss_mask"].sum()) == 0.0
    assert float(ds[0]["loss_mask"].sum()) > 0
    s = ResumableBatchSampler(10, 0, 4, dp_rank=1, dp_size=2, drop_last=False)
    s.pad_to_global_batch = True
    batches = list(s)
    assert batches == [[2, 3], [6, 7], [-1, -1]] and all(len(b) == 2 for b in batches)
    s0 = ResumableBatchSampler(10, 0, 4, dp_rank=0, dp_size=2, drop_last=False)
    s0.pad_to_global_batch = True
    assert list(s0)[-1] == [8, 9]


def test_preprocess_tool_roundtrip(tmp_path, monkeypatch):
    """tools/preprocess_data.py: jsonl -> Megatron .bin/.idx (readable by GPTDataset) and -> HF arrow blocks."""
    import importlib.util
    import json as _json
    spec = importlib.util.spec_from_file_location("preprocess_data", os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools", "preprocess_data.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)

    class Tok:
        eos_token_id, vocab_size = 2, 200

        def encode(self, text, add_special_tokens=False):
            return [3 + (ord(c) % 100) for c in text]
    monkeypatch.setattr(tool, "load_tokenizer", lambda args: Tok())
    src = tmp_path / "corpus.jsonl"
    texts = ["hello world " * (i + 1) for i in range(12)]
    src.write_text("\n".join(_json.dumps({"text": t}) for t in texts))
    info = tool.main(["--input", str(src), "--format", "megatron", "--output-prefix", str(tmp_path / "out" / "corpus"), "--append-eod"])
    assert info["documents"] == 12 and info["tokens"] == sum(len(t) + 1 for t in texts)
    indexed = MMapIndexedDataset(str(tmp_path / "out" / "corpus"))
    assert len(indexed) == 12 and list(indexed[0]) == Tok().encode(texts[0]) + [2]
    ds = GPTDataset("train", str(tmp_path / "out" / "corpus"), np.arange(12, dtype=np.int32), indexed, num_samples=8, seq_length=16, seed=1)
    assert ds[0]["tokens"].shape[0] == 16
    datasets = pytest.importorskip("datasets")
    info = tool.main(["--input", str(src), "--format", "hf", "--output-dir", str(tmp_path / "arrow"), "--seq-length", "32"])
    arrow = datasets.load_from_disk(str(tmp_path / "arrow"))
    assert len(arrow) == info["sequences"] > 0 and len(arrow[0]["input_ids"]) == 32 and arrow[0]["labels"] == arrow[0]["input_ids"]


def test_resumable_sampler_second_epoch_with_remainder():
    """len(dataset) % global_batch != 0: an epoch consumes floor(total/gbs) global batches, the next epoch starts at
    offset 0 again (it used to yield 0 batches and silently end training), and a mid-epoch resume continues in place."""
    from neuronx_distributed_training_b200.data.base import ResumableBatchSampler
    total, gbs = 1000, 64
    per_epoch = total // gbs
    consumed = 0
    seen = []
    for epoch in range(3):
        s = ResumableBatchSampler(total, consumed, gbs, dp_rank=0, dp_size=1)
        assert len(s) == per_epoch
        batches = list(s)
        assert len(batches) == per_epoch
        assert batches[0] == list(range(gbs))
        consumed += len(batches) * gbs
        seen.append(batches)
    assert seen[0] == seen[1] == seen[2]
    # resume 5 batches into the second epoch
    consumed = per_epoch * gbs + 5 * gbs
    s = ResumableBatchSampler(total, consumed, gbs, dp_rank=1, dp_size=2)
    rest = list(s)
    assert len(rest) == len(s) == per_epoch - 5
    assert rest[0] == list(range(5 * gbs + 32, 5 * gbs + 64))
    # shuffled: epochs differ, resume reproduces the tail of the same epoch
    a = list(ResumableBatchSampler(total, per_epoch * gbs, gbs, 0, 1, shuffle=True, seed=3))
    b = list(ResumableBatchSampler(total, 0, gbs, 0, 1, shuffle=True, seed=3))
    c = list(ResumableBatchSampler(total, per_epoch * gbs + 3 * gbs, gbs, 0, 1, shuffle=True, seed=3))
    assert a != b and c == a[3:]
    # drop_last=False: the padded tail batch counts as one more global batch of the epoch
    s = ResumableBatchSampler(total, 0, gbs, 0, 1, drop_last=False)
    assert len(s) == len(list(s)) == per_epoch + 1
    s = ResumableBatchSampler(total, (per_epoch + 1) * gbs, gbs, 0, 1, drop_last=False)
    assert len(list(s)) == per_epoch + 1
