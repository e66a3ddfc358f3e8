"""MX block-scaled fp8 GEMM (tcgen05.mma.kind::mxf8f6f4.block_scale, 1x32 E8M0 scales in tensor memory): numerics against the
dequantise-and-matmul reference of the same quantised operands, then timing against the bf16 and the row-scaled fp8 kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from neuronx_distributed_training_b200 import ops
from neuronx_distributed_training_b200.ops import fp8
C = ops.C()
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-9)).item()
torch.manual_seed(0)
ok = True
for (M, N, K) in [(128, 256, 128), (128, 256, 512), (256, 512, 384), (1024, 776, 1024), (384, 128, 256), (8192, 4096, 4096)]:
    # per-block dynamic range: rows and 32-blocks with very different magnitudes
    x = (torch.randn(M, K, device="cuda") * torch.exp(torch.randn(M, K // 32, device="cuda")).repeat_interleave(32, 1)).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05 * torch.exp(0.5 * torch.randn(N, K // 32, device="cuda")).repeat_interleave(32, 1)).bfloat16()
    xq, xs = fp8.quantize_mx(x); wq, ws = fp8.quantize_mx(w)
    xp, wp = fp8.pack_mx_scales(xs), fp8.pack_mx_scales(ws)
    ref = fp8.gemm_mx_reference(xq, xs, wq, ws, torch.float32)
    out = fp8.gemm_mx(xq, xp, wq, wp); torch.cuda.synchronize()
    e_bf16 = rel(out, ref)
    acc = torch.ones(M, N, device="cuda")
    fp8.gemm_mx(xq, xp, wq, wp, out=acc, accumulate=True); torch.cuda.synchronize()
    e_acc = rel(acc - 1.0, ref)
    f32 = torch.empty(M, N, device="cuda"); fp8.gemm_mx(xq, xp, wq, wp, out=f32); torch.cuda.synchronize()
    e_f32 = rel(f32, ref)
    exact = x.float() @ w.float().t()
    print(f"{M}x{N}x{K}: vs MX reference bf16 {e_bf16:.2e} fp32 {e_f32:.2e} fp32-accumulate {e_acc:.2e}; vs unquantised product {rel(out, exact):.2e}", flush=True)
    ok = ok and e_bf16 < 5e-3 and e_f32 < 1e-4 and e_acc < 1e-4
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for (M, N, K) in [(8192, 6144, 4096), (8192, 28672, 4096), (8192, 4096, 14336)]:
    x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
    xq, xs = fp8.quantize_mx(x); wq, ws = fp8.quantize_mx(w); xp, wp = fp8.pack_mx_scales(xs), fp8.pack_mx_scales(ws)
    rq, rs = C.fp8_quantize_rows(x); rwq, rws = C.fp8_quantize_rows(w)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    tm = t(lambda: C.gemm_mxfp8(xq, xp, wq, wp, out, False)); tr = t(lambda: C.gemm_fp8(rq, rs, rwq, rws, out, False))
    tb = t(lambda: C.gemm(x, w, out, True, True, False)); fl = 2.0 * M * N * K
    print(f"{M}x{N}x{K}: MX fp8 {tm:.3f} ms ({fl / tm / 1e9:.0f} TFLOP/s) | row-scaled fp8 {tr:.3f} ms ({fl / tr / 1e9:.0f}) | bf16 {tb:.3f} ms ({fl / tb / 1e9:.0f}) | MX / bf16 {tb / tm:.2f}x", flush=True)
print("MXFP8_OK" if ok else "MXFP8_MISMATCH")
