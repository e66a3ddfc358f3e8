import math
from neuronx_distributed_training_b200.optim.lr_schedulers import (LinearAnnealingWithWarmUp, CosineAnnealing, WarmupAnnealing,
                                                                   NoamAnnealing, get_scheduler, prepare_lr_scheduler)


class _Opt:
    def __init__(self, lr):
        self.param_groups = [{"lr": lr}]


def test_linear_warmup_decay():
    o = _Opt(1.0)
    s = LinearAnnealingWithWarmUp(o, max_steps=100, warmup_steps=10)
    lrs = []
    for _ in range(100):
        lrs.append(o.param_groups[0]["lr"]); s.step()
    assert lrs[0] == 0.0 and abs(lrs[5] - 0.5) < 1e-9 and abs(lrs[10] - 1.0) < 1e-9
    assert abs(lrs[55] - 0.5) < 1e-9 and o.param_groups[0]["lr"] == 0.0


def test_cosine_with_constant_tail():
    o = _Opt(2.0)
    s = CosineAnnealing(o, max_steps=110, warmup_steps=10, constant_steps=20, min_lr=0.2)
    vals = {}
    for i in range(111):
        vals[i] = o.param_groups[0]["lr"]; s.step()
    assert vals[0] < vals[5] < vals[10]
    mid = 10 + 40
    assert abs(vals[mid] - (0.2 + 0.5 * 1.8 * (1 + math.cos(math.pi * 40 / 80)))) < 1e-9
    assert abs(vals[95] - 0.2) < 1e-12 and abs(vals[110] - 0.2) < 1e-12


def test_state_dict_roundtrip_and_registry():
    o = _Opt(1.0)
    s = get_scheduler("WarmupAnnealing")(o, max_steps=50, warmup_steps=5)
    for _ in range(17):
        s.step()
    sd = s.state_dict()
    o2 = _Opt(1.0)
    s2 = WarmupAnnealing(o2, max_steps=50, warmup_steps=5)
    s2.load_state_dict(sd)
    assert o2.param_groups[0]["lr"] == o.param_groups[0]["lr"]
    s3 = prepare_lr_scheduler(_Opt(1.0), {"name": "NoamAnnealing", "warmup_steps": 10, "d_model": 1.0}, max_steps=100)
    assert isinstance(s3, NoamAnnealing)


def test_linear_annealing_with_warmup_matches_transformers_schedule():
    """The reference defines LinearAnnealingWithWarmUp as HF's linear-warmup / linear-decay LambdaLR (reference
    optim/lr_schedulers.py:11-23): compare step by step with transformers.get_linear_schedule_with_warmup."""
    import pytest
    import torch
    tr = pytest.importorskip("transformers")
    from neuronx_distributed_training_b200.optim.lr_schedulers import LinearAnnealingWithWarmUp

    class _Opt:
        def __init__(self):
            self.param_groups = [{"lr": 3e-4}]

    ours_opt = _Opt()
    ours = LinearAnnealingWithWarmUp(ours_opt, warmup_steps=5, max_steps=40)
    p = torch.nn.Parameter(torch.zeros(1))
    ref_opt = torch.optim.SGD([p], lr=3e-4)
    ref = tr.get_linear_schedule_with_warmup(ref_opt, num_warmup_steps=5, num_training_steps=40)
    for step in range(40):
        assert abs(ours_opt.param_groups[0]["lr"] - ref_opt.param_groups[0]["lr"]) < 1e-12, step
        ref_opt.step(); ref.step(); ours.step()
