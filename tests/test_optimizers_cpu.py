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
        get_optimizer("lamb_8bit")


def test_user_registered_update_rule_runs_inside_the_flat_optimizer():
    """optim.register_optimizer(name, callable): a custom element-wise rule (here: Lion) gets ZeRO-style flat buffers,
    clipping and state checkpointing from DistributedOptimizer and matches a hand-written reference."""
    from neuronx_distributed_training_b200 import optim
    from neuronx_distributed_training_b200.optim.distributed import DistributedOptimizer

    def lion(p32, g32, state, group, step):
        b1, b2 = group["betas"]
        m = state["exp_avg"]
        upd = (m * b1 + g32 * (1 - b1)).sign_()
        p32.mul_(1 - group["lr"] * group["weight_decay"]).add_(upd, alpha=-group["lr"])
        m.mul_(b2).add_(g32, alpha=1 - b2)

    optim.register_optimizer("lion", lion)
    algo, _ = optim.get_optimizer("lion")
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(7, 5)); b = torch.nn.Parameter(torch.randn(5))
    ref_w, ref_b = w.detach().clone(), b.detach().clone()
    ref_m = [torch.zeros_like(ref_w), torch.zeros_like(ref_b)]
    opt = DistributedOptimizer([{"params": [w, b]}], lr=1e-2, betas=(0.9, 0.99), weight_decay=0.1, algorithm=algo,
                               zero1=False, master_weights=False, fp32_grad_acc=False, max_grad_norm=None)
    for step in range(4):
        gs = [torch.randn(7, 5), torch.randn(5)]
        for p_, g_ in zip(opt.params, gs):
            if getattr(p_, "main_grad", None) is not None:
                p_.main_grad.copy_(g_)
            else:
                p_.grad = g_.clone()
        opt.step()
        for r, m, g_ in zip((ref_w, ref_b), ref_m, gs):
            upd = (m * 0.9 + g_ * 0.1).sign()
            r.mul_(1 - 1e-2 * 0.1).add_(upd, alpha=-1e-2)
            m.mul_(0.99).add_(g_, alpha=0.01)
    assert torch.allclose(opt.params[0].detach().view(7, 5), ref_w, atol=1e-6)
    assert torch.allclose(opt.params[1].detach(), ref_b, atol=1e-6)
    sd = opt.state_dict()
    assert sd["algorithm"] == "lion"


@pytest.mark.parametrize("beta1,scale_parameter,wd", [(None, True, 0.0), (0.9, False, 0.05), (None, False, 0.0)])
def test_adafactor_matches_the_reference_implementation(beta1, scale_parameter, wd):
    """Factored second moments inside the flat optimizer == transformers.optimization.Adafactor tensor by tensor
    (2-D, 3-D "stacked experts" and 1-D parameters; update clipping, parameter scaling, momentum, weight decay)."""
    hf_optim = pytest.importorskip("transformers.optimization")
    from neuronx_distributed_training_b200.optim.distributed import DistributedOptimizer
    from neuronx_distributed_training_b200 import optim
    torch.manual_seed(0)
    shapes = [(12, 7), (3, 5, 4), (9,), (6, 6)]
    ours = [torch.nn.Parameter(torch.randn(*s)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    algo, _ = optim.get_optimizer("adafactor")
    opt = DistributedOptimizer([{"params": ours}], lr=1e-2, weight_decay=wd, algorithm=algo, zero1=False, master_weights=False,
                               fp32_grad_acc=False, max_grad_norm=None, eps=(1e-30, 1e-3),
                               extra_defaults={"beta1": beta1, "scale_parameter": scale_parameter, "clip_threshold": 1.0,
                                               "decay_rate": -0.8})
    ropt = hf_optim.Adafactor(ref, lr=1e-2, eps=(1e-30, 1e-3), clip_threshold=1.0, decay_rate=-0.8, beta1=beta1,
                              weight_decay=wd, scale_parameter=scale_parameter, relative_step=False, warmup_init=False)
    for step in range(5):
        for p_, r_ in zip(opt.params, ref):
            g = torch.randn(*r_.shape) * (3.0 if step == 2 else 0.3)          # step 2 triggers the RMS clipping
            r_.grad = g.clone()
            p_.main_grad.copy_(g)
        opt.step()
        ropt.step()
        for p_, r_ in zip(opt.params, ref):
            assert torch.allclose(p_.detach(), r_.detach(), atol=2e-6, rtol=1e-5), (step, tuple(r_.shape))
    sd = opt.state_dict()
    assert "adafactor_rc" in sd
