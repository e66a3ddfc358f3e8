"""NVLS (NVSwitch multicast, multimem.*) kernels against NCCL + the plain RMSNorm kernel.
Run: torchrun --nproc-per-node N tests/gpu_checks/check_nvls.py   (prints NVLS_OK, or NVLS_UNAVAILABLE when the box maps no
multicast memory for symmetric allocations)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from neuronx_distributed_training_b200 import ops
from neuronx_distributed_training_b200.parallel import nvls
C = ops.C()
m, H = 8192 // world, 4096
try:
    ws = nvls.NvlsWorkspace(None, gather_bytes=world * m * H * 2)
except Exception as e:  # noqa: BLE001
    print(f"[rank {rank}] NVLS workspace: {type(e).__name__}: {e}", flush=True)
    dist.barrier()
    if rank == 0:
        print("NVLS_UNAVAILABLE", flush=True)
    dist.destroy_process_group(); sys.exit(0)
print(f"[rank {rank}] multicast_ptr={ws.mc_ptr:#x}", flush=True)
torch.manual_seed(1234 + rank)
ok = True
# 1. small all-reduce
for n in (4, 1000, 65536):
    t = torch.randn(n, device="cuda"); ref = t.clone(); dist.all_reduce(ref)
    for _ in range(3):                      # repeated: the two staging regions and the two counters alternate
        got = ws.all_reduce(t)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    ok = ok and err < 1e-5
    if rank == 0: print(f"all_reduce_ n={n}: max rel err {err:.2e}", flush=True)
# 2. RMSNorm + all-gather, repeated (double buffering, counters)
w = (1 + 0.1 * torch.randn(H, device="cuda")).bfloat16(); dist.broadcast(w, 0)
for it in range(5):
    x = torch.randn(m, H, device="cuda").bfloat16()
    y_ref, rstd_ref = C.rmsnorm_fwd(x, w, 1e-5)
    g_ref = torch.empty(world * m, H, device="cuda", dtype=torch.bfloat16); dist.all_gather_into_tensor(g_ref, y_ref)
    y, rstd = ws.rmsnorm_all_gather(x, w, 1e-5); torch.cuda.synchronize()
    # the row sum of squares is accumulated in a different order than in rmsnorm_fwd: rstd agrees to fp32 rounding and
    # the bf16 outputs to one ulp (2^-8 relative) on the few elements that sit on a rounding boundary
    d = (y.float() - g_ref.float()).abs(); e2 = (rstd - rstd_ref).abs().max().item()
    worst = (d / g_ref.float().abs().clamp_min(1e-3)).max().item(); frac = (d > 0).float().mean().item()
    ok = ok and worst <= 2 ** -6 and frac < 0.02 and e2 < 2e-6     # one bf16 ulp is at most 2^-7 relative
    if rank == 0: print(f"rmsnorm_all_gather it={it}: worst rel diff {worst:.2e} (1 bf16 ulp = 3.9e-3), differing elements {100*frac:.3f} %, rstd err {e2:.1e}", flush=True)
# 3. autograd wrapper vs norm + gather + autograd
x = torch.randn(m, 1, H, device="cuda").bfloat16().requires_grad_(True); wp = w.clone().requires_grad_(True)
out = nvls.rmsnorm_all_gather(x, wp, 1e-5, ws, keep=True)
gy = torch.randn_like(out); dist.broadcast(gy, 0)
out.backward(gy * (rank + 1))
x2 = x.detach().clone().requires_grad_(True); w2 = w.clone().requires_grad_(True)
y2 = ops.rmsnorm(x2, w2, 1e-5)
full = torch.empty(world * m, 1, H, device="cuda", dtype=torch.bfloat16); dist.all_gather_into_tensor(full, y2.detach())
# d(loss)/d(y_local) = sum over ranks of their gradient on my rows
gsum = gy.float() * sum(r + 1 for r in range(world))
y2.backward(gsum[rank * m:(rank + 1) * m].to(torch.bfloat16))
ex = ((x.grad.float() - x2.grad.float()).norm() / x2.grad.float().norm()).item()
ew = ((wp.grad.float() - w2.grad.float()).norm() / w2.grad.float().norm()).item()
ok = ok and ex < 2e-2 and ew < 2e-2
if rank == 0: print(f"autograd: dx rel err {ex:.2e} dw rel err {ew:.2e}", flush=True)
# 4. timing: fused norm+multicast vs norm kernel + NCCL all-gather
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return t.item()
x = torch.randn(m, H, device="cuda").bfloat16()
g = torch.empty(world * m, H, device="cuda", dtype=torch.bfloat16)
def nccl():
    y, _ = C.rmsnorm_fwd(x, w, 1e-5); dist.all_gather_into_tensor(g, y)
t_nvls = timeit(lambda: ws.rmsnorm_all_gather(x, w, 1e-5)); t_nccl = timeit(nccl)
s = torch.randn(8192, device="cuda")
t_ar = timeit(lambda: ws.all_reduce(s)); t_ar_nccl = timeit(lambda: dist.all_reduce(s))
if rank == 0:
    print(f"rmsnorm+all-gather [{world}x{m}x{H}] bf16: NVLS kernel {t_nvls*1e3:.1f} us | rmsnorm kernel + NCCL all_gather {t_nccl*1e3:.1f} us", flush=True)
    print(f"all-reduce 32 KB fp32: NVLS one-shot kernel {t_ar*1e3:.1f} us | NCCL {t_ar_nccl*1e3:.1f} us", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("NVLS_OK" if flag.item() == 1 else "NVLS_FAILED", flush=True)
dist.destroy_process_group()
