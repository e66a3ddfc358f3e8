"""fp8 (e4m3, row / column block scales) recipe on CPU through the PyTorch reference implementation."""
import pytest
import torch

from neuronx_distributed_training_b200 import ops
from neuronx_distributed_training_b200.ops import fp8
from neuronx_distributed_training_b200.config import from_dict, process_config
from neuronx_distributed_training_b200.parallel import state as ps


@pytest.fixture(autouse=True)
def _mp():
    ps.initialize_model_parallel(1, 1, 1)
    yield
    fp8.set_enabled(False)
    fp8.clear_cache()
    ps.destroy_model_parallel()


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_quantisation_round_trip():
    torch.manual_seed(0)
    x = torch.randn(64, 256) * torch.logspace(-2, 2, 64)[:, None]          # rows of very different magnitude
    q, s = fp8.quantize_rows(x)
    assert q.dtype == torch.uint8 and q.shape == x.shape and s.shape == (64,)
    assert _rel(fp8.dequantize(q, s), x) < 0.04                            # e4m3: 3 mantissa bits
    qt, st = fp8.quantize_cols_transposed(x)
    assert qt.shape == (256, 64) and st.shape == (256,)
    assert _rel(fp8.dequantize(qt, st).t(), x) < 0.04
    z = torch.zeros(4, 16)
    qz, sz = fp8.quantize_rows(z)
    assert torch.all(sz == 1) and torch.all(fp8.dequantize(qz, sz) == 0)


def test_three_gemms_of_a_linear_layer():
    torch.manual_seed(1)
    M, K, N = 96, 128, 64
    x, w, dy = torch.randn(M, K), torch.randn(N, K) * 0.1, torch.randn(M, N)
    assert _rel(fp8.linear_fwd(x, w), x @ w.t()) < 0.06
    assert _rel(fp8.linear_dgrad(dy, w), dy @ w) < 0.06
    acc = torch.ones(N, K)
    fp8.linear_wgrad(dy, x, out=acc, accumulate=True)
    assert _rel(acc - 1.0, dy.t() @ x) < 0.06
    # weight copies are cached until the parameter is modified in place
    n = len(fp8._WCACHE)
    fp8.linear_fwd(x, w)
    assert len(fp8._WCACHE) == n
    w.add_(1.0)
    assert _rel(fp8.linear_fwd(x, w), x @ w.t()) < 0.06                    # refreshed, not stale


def test_parallel_linear_uses_fp8_when_enabled_and_lm_head_does_not():
    from neuronx_distributed_training_b200.parallel.layers import ColumnParallelLinear
    torch.manual_seed(2)
    lin = ColumnParallelLinear(64, 32, bias=False, dtype=torch.float32)
    x = torch.randn(5, 3, 64, requires_grad=True)
    ref = lin(x)
    fp8.set_enabled(True)
    y = lin(x)
    y.sum().backward()
    assert 1e-4 < _rel(y, ref) < 0.06                                      # quantised, but close
    assert _rel(x.grad.reshape(15, 64), torch.ones(15, 32) @ lin.weight.detach()) < 0.06
    a, b = torch.randn(8, 64), torch.randn(16, 64)
    assert torch.equal(ops.gemm_nt(a, b, fp8=False), a @ b.t())            # explicit opt-out (LM head / CE)


def _module(tmp_path, tag):
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    cfg = process_config(from_dict({
        "name": "t", "model_source": "hf", "seed": 3,
        "trainer": {"devices": 1, "num_nodes": 1, "max_steps": 10, "gradient_clip_val": 1.0},
        "exp_manager": {"explicit_log_dir": str(tmp_path / tag), "create_tensorboard_logger": False, "create_checkpoint_callback": False},
        "distributed_strategy": {"tensor_model_parallel_size": 1, "zero1": True},
        "data": {"micro_batch_size": 1, "global_batch_size": 2, "synthetic": True},
        "model": {"model_config": "llama_125m", "encoder_seq_length": 32, "num_layers": 2, "hidden_size": 64, "intermediate_size": 128,
                  "num_attention_heads": 4, "num_kv_heads": 2, "vocab_size": 128,
                  "fusions": {"flash_attention": False, **({"fp8_gemm": True} if tag == "fp8" else {})},
                  "optim": {"name": "adamw_fp32OptState", "lr": 1e-2, "weight_decay": 0.0, "betas": [0.9, 0.95],
                            "sched": {"name": "LinearAnnealingWithWarmUp", "warmup_steps": 1, "max_steps": 50}}},
        "precision": {"type": "fp32"}}), 1)
    fp8.set_enabled(False)
    mod = HFLlamaModule(cfg, None)
    mod.setup()
    mod.configure_optimizers(50)
    return mod


def test_small_llama_in_fp8_gradients_track_the_unquantised_model_and_it_trains(tmp_path):
    ids = torch.randint(0, 128, (2, 32), generator=torch.Generator().manual_seed(0))
    batch = {"input_ids": ids, "labels": ids.clone()}
    grads, first_loss = {}, {}
    for tag in ("ref", "fp8"):
        mod = _module(tmp_path, tag)
        assert fp8.enabled() == (tag == "fp8")                     # switched on by model.fusions.fp8_gemm
        mod.optimizer.zero_grad()
        first_loss[tag] = float(mod.forward_backward_step(dict(batch)))
        grads[tag] = {n: p.main_grad.clone() for n, p in mod.local_named_parameters()}
        if tag == "fp8":
            losses = [float(mod.training_step(dict(batch))) for _ in range(10)]
            assert losses[-1] < losses[0] - 1.0, losses              # memorises the batch
    assert abs(first_loss["fp8"] - first_loss["ref"]) < 0.02
    for n, g in grads["ref"].items():
        assert _rel(grads["fp8"][n], g) < 0.12, (n, _rel(grads["fp8"][n], g))    # e4m3 has 3 mantissa bits
        assert abs(grads["fp8"][n].norm() / g.norm() - 1.0) < 0.05, n


def test_mx_block_scaled_reference_and_scale_packing():
    torch.manual_seed(3)
    x = torch.randn(200, 256) * torch.logspace(-3, 3, 256)[None, :]       # magnitudes vary along K: per-32 blocks matter
    q, s = fp8.quantize_mx(x)
    assert q.shape == (200, 256) and s.shape == (200, 8) and s.dtype == torch.uint8
    assert _rel(fp8.dequantize_mx(q, s), x) < 0.04
    qr, sr = fp8.quantize_rows(x)                                           # one scale per row loses the small columns
    small = slice(0, 64)
    assert _rel(fp8.dequantize_mx(q, s)[:, small], x[:, small]) < 0.05 < _rel(fp8.dequantize(qr, sr)[:, small], x[:, small])
    w = torch.randn(96, 256) * 0.1
    wq, ws = fp8.quantize_mx(w)
    assert _rel(fp8.gemm_mx_reference(q, s, wq, ws, torch.float32), x @ w.t()) < 0.05
    # hardware scale layout: 128 rows x 4 k-blocks per 512-byte tile, offset (r % 32) * 16 + (r // 32) * 4 + kb % 4
    packed = fp8.pack_mx_scales(s)
    assert packed.shape == (2, 2, 512)
    for (r, kb) in [(0, 0), (5, 3), (37, 1), (127, 7), (128, 4), (199, 6)]:
        tile = packed[r // 128, kb // 4]
        rr = r % 128
        assert int(tile[(rr % 32) * 16 + (rr // 32) * 4 + kb % 4]) == int(s[r, kb])
    assert torch.equal(fp8.unpack_mx_scales(packed, 200, 8), s)
    assert int(packed[1, 0, (100 % 32) * 16 + (100 // 32) * 4]) == 127     # row 228 does not exist: neutral scale


def test_mx_recipe_linear_layer_gemms_against_fp32():
    """recipe "mx" (1 x 32 E8M0 blocks): forward, dgrad and wgrad of a linear layer through the MX path (reference
    implementation on the CPU, the block-scaled tensor-core kernel on the GPU) stay within fp8 error of the fp32 products,
    and the packed scale tiles round-trip."""
    import torch
    from neuronx_distributed_training_b200.ops import fp8
    torch.manual_seed(0)
    M, N, K = 256, 384, 128
    x = torch.randn(M, K) * torch.exp(torch.randn(M, 1))
    w = torch.randn(N, K) * 0.05
    dy = torch.randn(M, N) * 0.1
    prev = fp8.recipe()
    fp8.set_recipe("mx")
    try:
        assert fp8._mx_ok(M, K, N)
        y = fp8.linear_fwd(x, w)
        dx = fp8.linear_dgrad(dy, w)
        dw = fp8.linear_wgrad(dy, x)
    finally:
        fp8.set_recipe(prev)
    def rel(a, b): return ((a.float() - b).norm() / b.norm()).item()
    assert rel(y, x @ w.t()) < 0.05 and rel(dx, dy @ w) < 0.05 and rel(dw, dy.t() @ x) < 0.05
    q, s = fp8.quantize_mx(x)
    assert torch.equal(fp8.unpack_mx_scales(fp8.pack_mx_scales(s), M, K // 32), s)
    # every block is representable: amax / 2^e <= 448, and the scale is the smallest such power of two
    xb = x.reshape(M, K // 32, 32).abs().amax(-1)
    e = s.float() - 127
    assert (xb / torch.exp2(e) <= 448.0 + 1e-3).all() and ((xb / torch.exp2(e - 1) > 448.0) | (xb == 0)).all()
