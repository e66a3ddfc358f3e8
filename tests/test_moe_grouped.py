"""Grouped (128-row padded) expert path vs the per-expert loop; layout helper invariants (CPU, torch reference ops)."""
import pytest
import torch

from neuronx_distributed_training_b200 import ops
from neuronx_distributed_training_b200.parallel import state as ps


@pytest.fixture(autouse=True)
def _mp():
    ps.initialize_model_parallel(1, 1, 1)
    yield
    ps.destroy_model_parallel()


def test_grouped_layout_invariants():
    counts = torch.tensor([5, 0, 130, 128, 1])
    n = int(counts.sum())
    dest, block_expert, offsets, t_cap = ops.grouped_layout(counts, n)
    assert t_cap % 128 == 0 and t_cap >= n
    assert offsets.tolist() == [0, 128, 128, 384, 512, 640]
    assert block_expert.tolist()[:5] == [0, 2, 2, 3, 4] and all(b == -1 for b in block_expert.tolist()[5:])
    assert dest.unique().numel() == n                                    # injective
    e_row = torch.repeat_interleave(torch.arange(5), counts)
    assert torch.all(dest >= offsets[e_row]) and torch.all(dest < offsets[e_row + 1])
    # order inside an expert is preserved
    assert dest[:5].tolist() == [0, 1, 2, 3, 4] and dest[5].item() == 128


@pytest.mark.parametrize("counts", [[7, 0, 33, 2], [0, 0, 5, 0], [128, 129, 1, 0]])
def test_grouped_experts_match_loop(counts, monkeypatch):
    from neuronx_distributed_training_b200.models.moe import ExpertMLPs
    torch.manual_seed(0)
    E, H, I = len(counts), 32, 48
    mlp = ExpertMLPs(E, H, I, glu=True, dtype=torch.float32, init_std=0.2)
    n = sum(counts)
    x = torch.randn(n, H, requires_grad=True)
    g = torch.randn(n, H)
    res = {}
    for mode in ("off", "on"):
        monkeypatch.setenv("NXDT_MOE_GROUPED", mode)
        for p in mlp.parameters():
            p.grad = None
        x.grad = None
        y = mlp.forward_sorted(x, torch.tensor(counts))
        (y * g).sum().backward()
        res[mode] = (y.detach().clone(), x.grad.clone(), mlp.mlp_op.gate_up_proj.grad.clone(), mlp.mlp_op.down_proj.grad.clone())
    for a, b in zip(res["off"], res["on"]):
        assert torch.allclose(a, b, atol=1e-5, rtol=1e-5)


def test_grouped_wgrad_accumulates_into_main_grad():
    torch.manual_seed(1)
    counts = torch.tensor([3, 0, 4])
    dest, be, offs, t_cap = ops.grouped_layout(counts, 7)
    x = torch.zeros(t_cap, 16).index_copy(0, dest, torch.randn(7, 16)).requires_grad_(True)
    w = torch.nn.Parameter(torch.randn(3, 16, 8))
    w.main_grad = torch.ones(3, 16, 8)
    y = ops.grouped_linear(x, w, be, offs)
    y.sum().backward()
    assert w.grad is None and w.grad_added_to_main_grad
    ref = torch.stack([x[offs[e]:offs[e + 1]].detach().t() @ torch.ones(int(offs[e + 1] - offs[e]), 8) for e in range(3)])
    assert torch.allclose(w.main_grad, 1.0 + ref, atol=1e-5)
