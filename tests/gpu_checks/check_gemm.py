"""Standalone GEMM correctness + speed probe (run under `timeout` so a hang cannot hold the box)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from neuronx_distributed_training_b200 import ops
C = ops.C()
assert C is not None, ops._native_loader.load_error()
torch.manual_seed(0)
res = {}
# GEMM_IMPL=2cta probes the SM-pair (cta_group::2) kernel with the same shapes
GEMM = C.gemm_2cta if os.environ.get("GEMM_IMPL") == "2cta" else C.gemm
TAG = "2cta" if os.environ.get("GEMM_IMPL") == "2cta" else "1cta"
def check(M, N, K, layout):
    if layout == "nt":
        a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16(); ref = a.float() @ b.float().t(); fl = (True, True)
    elif layout == "nn":
        a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(K, N, device="cuda").bfloat16(); ref = a.float() @ b.float(); fl = (True, False)
    else:
        a = torch.randn(K, M, device="cuda").bfloat16(); b = torch.randn(K, N, device="cuda").bfloat16(); ref = a.float().t() @ b.float(); fl = (False, False)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    GEMM(a, b, out, fl[0], fl[1], False); torch.cuda.synchronize()
    err = ((out.float() - ref).norm() / ref.norm()).item()
    print(f"gemm {layout} {M}x{N}x{K}: rel_err={err:.3e} finite={bool(torch.isfinite(out.float()).all())}", flush=True)
    return err
def check_acc(M, N, K, f32):
    a = torch.randn(K, M, device="cuda").bfloat16(); b = torch.randn(K, N, device="cuda").bfloat16()
    base = torch.randn(M, N, device="cuda")
    out = base.clone() if f32 else base.bfloat16()
    ref = out.float() + a.float().t() @ b.float()
    GEMM(a, b, out, False, False, True); torch.cuda.synchronize()
    err = ((out.float() - ref).norm() / ref.norm()).item()
    print(f"gemm tn accumulate f32={f32} {M}x{N}x{K}: rel_err={err:.3e}", flush=True)
    return err
res["acc_f32"] = check_acc(512, 768, 1024, True)
res["acc_bf16"] = check_acc(512, 768, 1024, False)
for layout in ("nt", "nn", "tn"):
    for (M, N, K) in [(128, 256, 64), (256, 256, 128), (1024, 768, 4096), (200, 264, 136)]:
        res[f"{layout}_{M}_{N}_{K}"] = check(M, N, K, layout)
def bench(M, N, K, layout, iters=20):
    fl = {"nt": (True, True), "nn": (True, False), "tn": (False, False), "tk": (False, True)}[layout]
    shp = {"nt": ((M, K), (N, K)), "nn": ((M, K), (K, N)), "tn": ((K, M), (K, N)), "tk": ((K, M), (N, K))}[layout]
    # rotate through enough distinct operand sets to defeat L2 residency (126 MB)
    nset = max(2, int(300e6 // ((M * K + N * K + M * N) * 2)) + 1)
    As = [torch.randn(*shp[0], device="cuda").bfloat16() for _ in range(nset)]
    Bs = [torch.randn(*shp[1], device="cuda").bfloat16() for _ in range(nset)]
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for i in range(3): GEMM(As[i % nset], Bs[i % nset], out, fl[0], fl[1], False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): GEMM(As[i % nset], Bs[i % nset], out, fl[0], fl[1], False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # cuBLAS on the same operands
    def tref(i):
        if layout == "nt": return As[i % nset] @ Bs[i % nset].t()
        if layout == "nn": return As[i % nset] @ Bs[i % nset]
        if layout == "tk": return As[i % nset].t() @ Bs[i % nset].t()
        return As[i % nset].t() @ Bs[i % nset]
    for i in range(3): tref(i)
    torch.cuda.synchronize(); e0.record()
    for i in range(iters): tref(i)
    e1.record(); torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9; tfr = 2.0 * M * N * K / ms_ref / 1e9
    print(f"bench[{TAG}] {layout} {M}x{N}x{K}: native {ms:.3f} ms {tf:.0f} TFLOP/s | cuBLAS {ms_ref:.3f} ms {tfr:.0f} TFLOP/s | ratio {tf/tfr:.2f}", flush=True)
    res[f"bench_{layout}_{M}_{N}_{K}"] = {"native_tflops": tf, "cublas_tflops": tfr}
for (M, N, K, l) in [(8192, 6144, 4096, "nt"), (8192, 28672, 4096, "nt"), (8192, 4096, 14336, "nt"),
                     (8192, 4096, 28672, "nn"), (28672, 4096, 8192, "tn"), (8192, 8192, 8192, "nt"),
                     (4096, 8192, 28672, "tk"), (4096, 8192, 6144, "tk"), (8192, 768, 4096, "nt"), (8192, 4096, 3584, "nt"), (3584, 4096, 8192, "tn")]:
    bench(M, N, K, l)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/check_gemm_{TAG}.json", "w"), indent=1)
