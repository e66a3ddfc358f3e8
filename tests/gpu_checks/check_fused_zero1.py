"""Fused ZeRO-1 (peer-memory RS + AdamW + AG) vs the NCCL path of DistributedOptimizer (torchrun, DP=N)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from neuronx_distributed_training_b200.parallel import state as ps
from neuronx_distributed_training_b200.optim import DistributedOptimizer
from neuronx_distributed_training_b200.optim.fused_zero1 import enable_fused_zero1
ps.init_distributed("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
ps.initialize_model_parallel(1, 1, 1)
dev = torch.device("cuda", torch.cuda.current_device())
def make():
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(1024, 4096), torch.nn.GELU(), torch.nn.Linear(4096, 1024)).to(dev).bfloat16()
    decay = [p for n, p in m.named_parameters() if not n.endswith("bias")]; nod = [p for n, p in m.named_parameters() if n.endswith("bias")]
    opt = DistributedOptimizer([{"params": decay, "weight_decay": 0.1}, {"params": nod, "weight_decay": 0.0}], lr=1e-2, betas=(0.9, 0.95),
                               zero1=True, master_weights=True, fp32_grad_acc=True, max_grad_norm=0.5, dp_group=ps.get_data_parallel_group())
    return m, opt
m1, o1 = make(); m2, o2 = make()
assert enable_fused_zero1(o2)
g = torch.Generator(device="cpu").manual_seed(100 + rank)
for step in range(4):
    x = torch.randn(64, 1024, generator=g).to(dev).bfloat16(); y = torch.randn(64, 1024, generator=g).to(dev).bfloat16()
    # start every step from identical state so the comparison is a one-step equivalence (Adam amplifies ulp-level
    # differences of near-zero gradients, so trajectories may drift apart legitimately)
    with torch.no_grad():
        o2.param_buffer.copy_(o1.param_buffer); o2.master.copy_(o1.master)
        o2.exp_avg.copy_(o1.exp_avg); o2.exp_avg_sq.copy_(o1.exp_avg_sq)
    for m, o in ((m1, o1), (m2, o2)):
        o.zero_grad()
        ((m(x) - y).float() ** 2).mean().backward()
        o.step()
    torch.cuda.synchronize()
    dn = abs(float(o1.grad_norm) - float(o2.grad_norm))
    err = max((p.float() - q.float()).abs().max().item() for p, q in zip(m1.parameters(), m2.parameters()))
    merr = (o1.master - o2.master).abs().max().item()
    if rank == 0: print(f"step {step}: grad_norm {float(o1.grad_norm):.5f} vs fused {float(o2.grad_norm):.5f}  param max|diff| {err:.3e} master {merr:.3e}", flush=True)
    assert dn < 1e-3 and err < 4e-3 and merr < 1e-6, (dn, err, merr)
# timing on a larger flat buffer
def bench(o, iters=5):
    for _ in range(2): o.step()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): o.step()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / iters
big = [torch.nn.Parameter(torch.randn(8192, 8192, device=dev).bfloat16()) for _ in range(4)]
for p in big: p.tensor_model_parallel = True
oa = DistributedOptimizer([{"params": big}], lr=1e-3, zero1=True, dp_group=ps.get_data_parallel_group())
t_nccl = bench(oa)
enable_fused_zero1(oa)
t_fused = bench(oa)
if rank == 0: print(f"ZERO1 step on 268M params, DP={world}: nccl path {t_nccl:.3f} ms, fused peer-memory path {t_fused:.3f} ms\nFUSED_ZERO1_OK", flush=True)
dist.barrier(); dist.destroy_process_group()
