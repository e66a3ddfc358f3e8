"""Fused AG->GEMM / GEMM->RS / AG-wgrad kernels vs NCCL + matmul (torchrun --nproc-per-node N)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from neuronx_distributed_training_b200 import ops
from neuronx_distributed_training_b200.parallel import state as ps
from neuronx_distributed_training_b200.parallel.fused import FusedTP

ps.init_distributed("nccl")
tp = dist.get_world_size(); rank = dist.get_rank()
ps.initialize_model_parallel(tensor_model_parallel_size=tp)
dev = torch.device("cuda", torch.cuda.current_device())
S = int(os.environ.get("SEQ", "8192")); H = 4096; I = 14336
m = S // tp
ftp = FusedTP(ps.get_tensor_model_parallel_group(), max_shard_bytes=m * H * 2, max_stage_bytes=S * H * 2)
torch.manual_seed(100 + rank)
res = {}
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-9)).item()
def ag(x):
    out = torch.empty(x.shape[0] * tp, x.shape[1], dtype=x.dtype, device=dev); dist.all_gather_into_tensor(out, x); return out
def rs(x):
    out = torch.empty(x.shape[0] // tp, x.shape[1], dtype=x.dtype, device=dev); dist.reduce_scatter_tensor(out, x); return out
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); return t.item()

# ---- AG -> GEMM (column fwd: qkv / gate_up) ----
for name, N in (("qkv", 6144 // tp), ("gate_up", 2 * I // tp)):
    x = (torch.randn(m, H, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, H, device=dev) * 0.05).bfloat16()
    y = ftp.all_gather_gemm(x, w); torch.cuda.synchronize()
    ref = ag(x).float() @ w.float().t()
    e = rel(y, ref); res[f"ag_gemm_{name}_err"] = e
    t_f = timeit(lambda: ftp.all_gather_gemm(x, w)); t_b = timeit(lambda: ag(x) @ w.t())
    res[f"ag_gemm_{name}_ms"] = {"fused": t_f, "nccl+cublas": t_b}
    if rank == 0: print(f"AG->GEMM {name}: N={N} rel_err={e:.2e} fused {t_f:.3f} ms vs nccl+cublas {t_b:.3f} ms", flush=True)
# ---- GEMM -> RS (row fwd: o_proj / down) ----
for name, K in (("o_proj", 4096 // tp), ("down", I // tp)):
    a = (torch.randn(S, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(H, K, device=dev) * 0.05).bfloat16()
    z = ftp.gemm_reduce_scatter(a, w); torch.cuda.synchronize()
    part = (a.float() @ w.float().t())
    dist.all_reduce(part); ref = part[rank * m:(rank + 1) * m]
    e = rel(z, ref); res[f"gemm_rs_{name}_err"] = e
    t_f = timeit(lambda: ftp.gemm_reduce_scatter(a, w)); t_b = timeit(lambda: rs(a @ w.t()))
    res[f"gemm_rs_{name}_ms"] = {"fused": t_f, "nccl+cublas": t_b}
    if rank == 0: print(f"GEMM->RS {name}: K={K} rel_err={e:.2e} fused {t_f:.3f} ms vs cublas+nccl {t_b:.3f} ms", flush=True)
# ---- dgrad conjugates (trans_b=False) ----
w = (torch.randn(2 * I // tp, H, device=dev) * 0.05).bfloat16()
dy = (torch.randn(S, 2 * I // tp, device=dev) * 0.5).bfloat16()
dx = ftp.gemm_reduce_scatter(dy, w, trans_b=False); torch.cuda.synchronize()
part = dy.float() @ w.float(); dist.all_reduce(part)
e = rel(dx, part[rank * m:(rank + 1) * m]); res["gemm_rs_dgrad_err"] = e
if rank == 0: print(f"GEMM->RS dgrad rel_err={e:.2e}", flush=True)
wr = (torch.randn(H, I // tp, device=dev) * 0.05).bfloat16(); dz = (torch.randn(m, H, device=dev) * 0.5).bfloat16()
da, dzg = ftp.all_gather_gemm(dz, wr, trans_b=False, return_gathered=True); torch.cuda.synchronize()
e = rel(da, ag(dz).float() @ wr.float()); e2 = rel(dzg, ag(dz)); res["ag_gemm_dgrad_err"] = e; res["ag_gathered_err"] = e2
if rank == 0: print(f"AG->GEMM dgrad rel_err={e:.2e} gathered_err={e2:.2e}", flush=True)
# ---- wgrad with the gather fused into the K loop ----
x = (torch.randn(m, H, device=dev) * 0.5).bfloat16()
mg = torch.zeros(2 * I // tp, H, device=dev)
ftp.all_gather_wgrad(dy, x, mg, True); ftp.all_gather_wgrad(dy, x, mg, True); torch.cuda.synchronize()
ref = 2 * (dy.float().t() @ ag(x).float())
e = rel(mg, ref); res["ag_wgrad_err"] = e
t_f = timeit(lambda: ftp.all_gather_wgrad(dy, x, mg, True)); t_b = timeit(lambda: mg.add_(dy.t() @ ag(x)))
res["ag_wgrad_ms"] = {"fused": t_f, "nccl+cublas": t_b}
if rank == 0: print(f"AG-wgrad rel_err={e:.2e} fused {t_f:.3f} ms vs nccl+cublas {t_b:.3f} ms", flush=True)
xg = ftp.all_gather(x); torch.cuda.synchronize()
res["ag_only_err"] = rel(xg, ag(x))
# ---- in-kernel timelines (globaltimer stamps per CTA) ----
if os.environ.get("FUSED_TIMELINE", "1") == "1":
    C = ops.C()
    dbg = torch.zeros(148, 4, dtype=torch.int64, device=dev)
    def timeline(name, fn):
        fn(); torch.cuda.synchronize(); dist.barrier()
        dbg.zero_(); C.fused_tp_set_debug(dbg)
        fn(); torch.cuda.synchronize()
        C.fused_tp_set_debug(torch.empty(0, dtype=torch.int64, device=dev))
        d = dbg.cpu()
        t0 = d[:, 0][d[:, 0] > 0].min().item()
        def stat(col):
            v = d[:, col][d[:, col] > 0]
            return (None if v.numel() == 0 else (round((v.min().item() - t0) / 1e3, 1), round((v.max().item() - t0) / 1e3, 1)))
        if rank == 0:
            print(f"TIMELINE {name}: start..{stat(0)} hs_done {stat(1)} main_done {stat(2)} end {stat(3)} (us, min/max over CTAs)", flush=True)
    x_ = (torch.randn(m, H, device=dev) * 0.5).bfloat16(); w_ = (torch.randn(6144 // tp, H, device=dev) * 0.05).bfloat16()
    timeline("ag_gemm_qkv", lambda: ftp.all_gather_gemm(x_, w_))
    a_ = (torch.randn(S, 4096 // tp, device=dev) * 0.5).bfloat16(); wo_ = (torch.randn(H, 4096 // tp, device=dev) * 0.05).bfloat16()
    timeline("gemm_rs_oproj", lambda: ftp.gemm_reduce_scatter(a_, wo_))
    timeline("ag_wgrad", lambda: ftp.all_gather_wgrad(dy, x, mg, True))
ok = all(v < 2e-2 for k, v in res.items() if k.endswith("_err"))
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/check_fused_tp_n{tp}.json", "w"), indent=1)
    print("FUSED_TP_OK" if ok else "FUSED_TP_MISMATCH", flush=True)
dist.barrier(); dist.destroy_process_group()
