from neuronx_distributed_training_b200.parallel.state import enumerate_groups, rank_coords, coords_to_rank
from neuronx_distributed_training_b200.pipeline.partition import even_partition, cuts_partition, stage_layer_ranges
import pytest


def test_megatron_layout_16_gpus():
    # the layout documented by the reference (models/megatron/megatron_init.py:104-119): TP=2, PP=4, 16 GPUs
    g = enumerate_groups(16, tp=2, pp=4, cp=1)
    assert [0, 1] in g["tp"] and [14, 15] in g["tp"] and len(g["tp"]) == 8
    assert [0, 4, 8, 12] in g["pp"] and [3, 7, 11, 15] in g["pp"] and len(g["pp"]) == 4
    assert [0, 2] in g["dp"] and [13, 15] in g["dp"] and len(g["dp"]) == 8


def test_every_rank_in_exactly_one_group_of_each_kind():
    for world, tp, pp, cp, ep in [(8, 2, 2, 1, 1), (8, 2, 1, 2, 1), (16, 2, 2, 2, 1), (8, 1, 1, 1, 4), (8, 2, 1, 1, 2)]:
        g = enumerate_groups(world, tp, pp, cp, ep)
        for kind, lists in g.items():
            if kind == "embed":
                continue
            seen = sorted(r for l in lists for r in l)
            assert seen == list(range(world)), (kind, world, tp, pp, cp, ep)


def test_coords_roundtrip():
    tp, cp, dp, pp = 2, 2, 2, 2
    for r in range(16):
        c = rank_coords(r, tp, cp, dp, pp)
        assert coords_to_rank(*c, tp, cp, dp) == r
    # TP is the fastest-varying dimension
    assert rank_coords(1, tp, cp, dp, pp) == (0, 0, 0, 1)
    assert rank_coords(2, tp, cp, dp, pp) == (0, 0, 1, 0)


def test_cp_tp_groups_and_ep():
    g = enumerate_groups(8, tp=2, pp=1, cp=2)
    assert [0, 2] in g["cp"] and [1, 3] in g["cp"]
    assert [0, 1, 2, 3] in g["tp_cp"]
    assert [0, 4] in g["dp"] and [0, 2, 4, 6] in g["dp_cp"]
    g = enumerate_groups(8, tp=1, pp=1, cp=1, ep=4)
    assert [0, 1, 2, 3] in g["ep"] and [4, 5, 6, 7] in g["ep"] and [0, 4] in g["edp"]


def test_partitions():
    assert even_partition(32, 4) == [(0, 8), (8, 16), (16, 24), (24, 32)]
    assert even_partition(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert cuts_partition(32, ["model.layers.7", "model.layers.15", "model.layers.23"]) == [(0, 8), (8, 16), (16, 24), (24, 32)]
    r = stage_layer_ranges(8, pp=2, vpp=2)
    assert r == [[(0, 2), (4, 6)], [(2, 4), (6, 8)]]
    with pytest.raises(ValueError):
        stage_layer_ranges(32, 4, 1, ["model.layers.7"])
