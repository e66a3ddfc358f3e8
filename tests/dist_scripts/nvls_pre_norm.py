"""The norm-inside-the-column-linear call path (parallel/nvls.py) with an emulated workspace: a sequence-parallel Llama
whose RMSNorm + all-gather pairs run through `_ColumnParallelLinearFn(..., norm_weight, eps, nvls)` must give the same loss
and the same gradient for every parameter as the ordinary norm -> linear path, with and without keeping the gathered input.
(The multicast kernel itself is validated on GPUs: tests/gpu_checks/check_nvls.py.)"""
from common import *
from neuronx_distributed_training_b200 import ops
from neuronx_distributed_training_b200.models.llama import LlamaConfig, LlamaForCausalLM
from neuronx_distributed_training_b200.parallel import layers as L
from neuronx_distributed_training_b200.parallel import mappings as mp

TP = int(os.environ.get("TP", "2"))
rank, world = init(tp=TP)


class EmulatedWorkspace:
    """rmsnorm_all_gather with torch ops + a gloo all-gather; counts calls."""
    def __init__(self):
        self.world, self.gather_bytes, self.calls = TP, 1 << 30, 0
        self.group = ps.get_tensor_model_parallel_group()

    def rmsnorm_all_gather(self, x2d, w, eps):
        self.calls += 1
        rstd = torch.rsqrt(x2d.float().pow(2).mean(-1) + eps)
        y = (x2d.float() * rstd.unsqueeze(-1) * w.float()).to(x2d.dtype)
        return mp.all_gather_dim0(y, self.group), rstd


g = torch.Generator().manual_seed(5)
ids = torch.randint(0, 64, (2, 8 * TP), generator=g)
base = dict(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
            num_key_value_heads=2, torch_dtype="float32", use_flash_attention=False, use_cpu_initialization=True,
            sequence_parallel_enabled=True)


def run(fuse_qkv, nvls, keep):
    torch.manual_seed(11)
    m = LlamaForCausalLM(LlamaConfig(**base, fuse_qkv=fuse_qkv))
    ws = None
    if nvls:
        ws = EmulatedWorkspace()
        for layer in m.model.layers.values() if hasattr(m.model.layers, "values") else m.model.layers:
            layer.nvls_pre_norm = True
            layer.self_attn.qkv_proj.nvls = ws
            layer.mlp.gate_up_proj.nvls = ws
    L.set_keep_gathered_input(keep)
    loss, _ = m(input_ids=ids, labels=ids)
    loss.backward()
    L.set_keep_gathered_input(False)
    return loss.detach().clone(), {n: p.grad.clone() for n, p in m.named_parameters()}, ws


# the emulation runs on CPU fp32: lift the CUDA / bf16 conditions of the selector, keep the rest of its logic
orig = L._nvls_for
L._nvls_for = lambda layer, x, pre_norm: getattr(layer, "nvls", None) if pre_norm is not None and layer.sequence_parallel_enabled else None
for fuse_qkv in (True, False):
    l0, g0, _ = run(fuse_qkv, False, False)
    for keep in (False, True):
        l1, g1, ws = run(fuse_qkv, True, keep)
        # forward: 2 layers x (qkv when it is one fused projection + gate_up); backward re-gathers when nothing was kept
        per_fwd = 2 * (2 if fuse_qkv else 1)
        assert ws.calls == (per_fwd if keep else 2 * per_fwd), (fuse_qkv, keep, ws.calls)
        check(f"loss fuse_qkv={fuse_qkv} keep={keep}", l1, l0, atol=1e-6)
        assert set(g0) == set(g1)
        for n in g0:
            check(f"grad {n} fuse_qkv={fuse_qkv} keep={keep}", g1[n], g0[n], atol=2e-6)
L._nvls_for = orig
finish()
