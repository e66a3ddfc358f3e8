"""Property-based tests (hypothesis) of host-side logic whose contract is easy to state and easy to get subtly wrong:
sequence packing, the resumable batch sampler, the YAML interpolation resolvers."""
import pytest
import torch

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st

from neuronx_distributed_training_b200.data.alignment import ConcatDataset, IGNORE_INDEX
from neuronx_distributed_training_b200.data.base import ResumableBatchSampler


@settings(max_examples=200, deadline=None)
@given(lengths=st.lists(st.integers(1, 40), min_size=1, max_size=30), seq=st.integers(8, 32))
def test_packing_preserves_tokens_order_and_row_width(lengths, seq):
    """ConcatDataset: every sample's (possibly truncated) tokens appear exactly once, in order, never split across rows; rows
    are exactly ``seq`` wide; padding carries attention 0 / label -100; the first label of every packed sample is masked."""
    samples, tok = [], 1
    for n in lengths:
        ids = list(range(tok, tok + n)); tok += n
        samples.append({"input_ids": ids, "attention_mask": [1] * n, "labels": list(ids)})
    ds = ConcatDataset([{k: list(v) for k, v in s.items()} for s in samples], seq, 0)
    flat_ids, starts = [], set()
    for i in range(len(ds)):
        row = ds[i]
        assert all(row[k].shape == (seq,) for k in ("input_ids", "attention_mask", "labels"))
        real = row["attention_mask"].bool()
        assert (row["input_ids"][~real] == 0).all() and (row["labels"][~real] == IGNORE_INDEX).all()
        assert real.int().diff().le(0).all()                      # real tokens first, padding after
        flat_ids.append(row["input_ids"][real].tolist())
        lab = row["labels"][real]
        ids = row["input_ids"][real]
        assert ((lab == IGNORE_INDEX) | (lab == ids)).all()
        starts |= {int(t) for t, l in zip(ids.tolist(), lab.tolist()) if l == IGNORE_INDEX}
    expect = [s["input_ids"][:seq] for s in samples]
    assert sum(flat_ids, []) == sum(expect, [])                   # nothing lost, duplicated or reordered
    assert all(e[0] in starts for e in expect)                    # each sample's first token is never a target
    # a sample is never split: each (truncated) sample is a contiguous slice of exactly one row
    for e in expect:
        assert any(_contains(r, e) for r in flat_ids)


def _contains(row, piece):
    n = len(piece)
    return any(row[i:i + n] == piece for i in range(len(row) - n + 1))


@settings(max_examples=150, deadline=None)
@given(total=st.integers(1, 200), per_rank=st.integers(1, 4), dp=st.integers(1, 4), steps=st.integers(0, 40),
       drop_last=st.booleans(), shuffle=st.booleans())
def test_sampler_resume_and_partition(total, per_rank, dp, steps, drop_last, shuffle):
    """ResumableBatchSampler: (a) the DP ranks' slices of a step tile that step's global batch without overlap, (b) a sampler
    resumed at ``consumed = steps * GBS`` continues exactly where the uninterrupted one is after ``steps`` batches -- across
    epoch boundaries too -- and (c) with drop_last an epoch never yields an index twice."""
    gbs = per_rank * dp
    if drop_last and total < gbs:
        return
    def batches(consumed, rank, n):
        out, c = [], consumed
        while len(out) < n:
            s = ResumableBatchSampler(total, c, gbs, rank, dp, drop_last=drop_last, shuffle=shuffle, seed=7)
            s.pad_to_global_batch = not drop_last
            got = list(s)
            assert len(got) == len(s)
            if not got:                                            # exactly at an epoch boundary
                raise AssertionError("a sampler positioned inside the schedule yielded nothing")
            out += got
            c += len(got) * gbs
        return out[:n]
    n = steps + 3
    ranks = [batches(0, r, n) for r in range(dp)]
    for step in range(n):
        glob = sum((ranks[r][step] for r in range(dp)), [])
        assert len(glob) == gbs
        real = [i for i in glob if i >= 0]
        assert len(set(real)) == len(real) and all(0 <= i < total for i in real)
    resumed = batches(steps * gbs, 0, 3)
    assert resumed == ranks[0][steps:steps + 3]
    if drop_last:
        per_epoch = total // gbs
        first_epoch = sum((sum((ranks[r][s] for r in range(dp)), []) for s in range(min(per_epoch, n))), [])
        assert len(set(first_epoch)) == len(first_epoch)


from neuronx_distributed_training_b200.config import from_dict
from neuronx_distributed_training_b200.config.loader import apply_overrides

_name = st.sampled_from(["alpha", "beta", "gamma", "delta"])


@settings(max_examples=150, deadline=None)
@given(a=st.integers(1, 64), b=st.integers(1, 64), c=st.integers(1, 8), key=_name)
def test_interpolation_and_multiply_resolver(a, b, c, key):
    """`${path}` references and the reference's `${multiply:x,y}` resolver, nested and chained, evaluate like Python."""
    cfg = from_dict({"data": {"micro": a, "dp": b, key: "${data.micro}"},
                     "trainer": {"global": "${multiply:${data.micro},${data.dp}}",
                                 "nested": "${multiply:${trainer.global},%d}" % c,
                                 "text": "run_${data.micro}_x_${data.dp}"}})
    assert cfg.data[key] == a
    assert cfg.trainer["global"] == a * b
    assert cfg.trainer.nested == a * b * c
    assert cfg.trainer.text == f"run_{a}_x_{b}"
    # command-line overrides are applied before interpolation is observed
    cfg2 = apply_overrides(cfg, [f"data.micro={a + 1}"])
    assert cfg2.trainer["global"] == (a + 1) * b and cfg2.trainer.nested == (a + 1) * b * c
