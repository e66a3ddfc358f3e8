"""Bandwidth of the flat AdamW step and of sq_norm (numerics are
covered by `pytest tests/test_kernels_gpu.py -k adamw_flat`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from neuronx_distributed_training_b200 import ops
n = 1 << 30                     # 1 Gi parameters: 30 GiB of traffic per step
master = torch.randn(n, device="cuda"); p = master.bfloat16()
g = torch.randn(n, device="cuda"); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
scale = torch.tensor(1.0, device="cuda")
def step(i):
    ops.adamw_step_flat(master, p, g, m, v, 1e-3, 0.9, 0.95, 1e-8, 0.01, i, grad_scale=scale)
for i in range(1, 4): step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(4, 9): step(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"adamw_flat: {ms:.3f} ms per 2^30 params, {30 * n / ms / 1e9:.2f} TB/s", flush=True)
for _ in range(3): ops.sq_norm(g)
torch.cuda.synchronize(); e0.record()
for _ in range(5): ops.sq_norm(g)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"sq_norm fp32: {ms:.3f} ms, {4 * n / ms / 1e9:.2f} TB/s", flush=True)
print("ADAM_CHECK_OK")
