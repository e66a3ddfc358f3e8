"""Flat-buffer optimizer algorithms vs torch.optim / a hand-written NovoGrad (single process, fp32)."""
import pytest
import torch

from neuronx_distributed_training_b200.optim import DistributedOptimizer, get_optimizer
from neuronx_distributed_training_b200.parallel import state as ps


@pytest.fixture(autouse=True)
def _mp():
    ps.initialize_model_parallel(1, 1, 1)
    yield
    ps.destroy_model_parallel()


def _models():
    torch.manual_seed(0)
    a = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
    b = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))
    b.load_state_dict(a.state_dict())
    return a, b


def _run(mine, ref_opt, model_a, model_b, steps=5):
    for i in range(steps):
        x = torch.randn(5, 8, generator=torch.Generator().manual_seed(i))
        for m, o in ((model_a, mine), (model_b, ref_opt)):
            o.zero_grad()
            m(x).pow(2).mean().backward()
            o.step()
    for pa, pb in zip(model_a.parameters(), model_b.parameters()):
        assert torch.allclose(pa, pb, atol=2e-6, rtol=1e-5), (pa - pb).abs().max()


@pytest.mark.parametrize("name,ref,kw", [
    ("adamax", torch.optim.Adamax, dict(lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.0)),
    ("adagrad", torch.optim.Adagrad, dict(lr=1e-2, eps=1e-10, weight_decay=0.0)),
    ("rmsprop", torch.optim.RMSprop, dict(lr=1e-3, eps=1e-8, weight_decay=0.0, alpha=0.99)),
    ("sgd", torch.optim.SGD, dict(lr=1e-2, weight_decay=0.0, momentum=0.9)),
])
def test_matches_torch_optim(name, ref, kw):
    a, b = _models()
    algo, extra = get_optimizer(name)
    mine_kw = {k: v for k, v in kw.items() if k != "alpha"}
    groups = [{"params": list(a.parameters()), **({"alpha": kw["alpha"]} if "alpha" in kw else {})}]
    mine = DistributedOptimizer(groups, algorithm=algo, zero1=False, max_grad_norm=None, dp_group="self", **mine_kw, **extra)
    _run(mine, ref(b.parameters(), **kw), a, b)


def test_novograd_layerwise_second_moment():
    a, b = _models()
    lr, b1, b2, eps, wd = 1e-2, 0.95, 0.98, 1e-8, 0.01
    mine = DistributedOptimizer([{"params": list(a.parameters())}], algorithm="novograd", lr=lr, betas=(b1, b2), eps=eps,
                                weight_decay=wd, zero1=False, max_grad_norm=None, dp_group="self")
    ms = [torch.zeros_like(p) for p in b.parameters()]
    vs = [None] * len(ms)
    for i in range(4):
        x = torch.randn(5, 8, generator=torch.Generator().manual_seed(i))
        mine.zero_grad()
        a(x).pow(2).mean().backward()
        mine.step()
        b.zero_grad()
        b(x).pow(2).mean().backward()
        with torch.no_grad():
            for j, p in enumerate(b.parameters()):
                n2 = p.grad.pow(2).sum()
                vs[j] = n2 if vs[j] is None else b2 * vs[j] + (1 - b2) * n2
                upd = p.grad / (vs[j].sqrt() + eps) + wd * p
                ms[j].mul_(b1).add_(upd)
                p.add_(ms[j], alpha=-lr)
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, atol=2e-6, rtol=1e-5)
    sd = mine.state_dict()
    assert sd["layer_v"].numel() == 4


def test_unknown_optimizer_name():
    with pytest.raises(KeyError):
        get_optimizer("adafactor")
