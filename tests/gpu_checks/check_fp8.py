"""Numerics + timing driver for the fp8 (e4m3) tcgen05 GEMM and its quantisation kernels (single GPU).

Compares the native kernels with the PyTorch float8 reference of the same recipe and with the fp32
product, then times the Llama-3-8B shapes against the bf16 kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from neuronx_distributed_training_b200 import ops  # noqa: E402
from neuronx_distributed_training_b200.ops import fp8  # noqa: E402


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-9)).item()


def numerics():
    C = ops.C()
    torch.manual_seed(0)
    for (M, N, K) in [(128, 256, 128), (256, 512, 384), (1000, 776, 1040), (8192, 4096, 4096)]:
        x = (torch.randn(M, K, device="cuda") * torch.rand(M, 1, device="cuda") * 4).bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        q, s = C.fp8_quantize_rows(x)
        fp8.use_native_kernels(False)
        q_ref, s_ref = fp8.quantize_rows(x)
        fp8.use_native_kernels(True)
        assert torch.allclose(s, s_ref, rtol=1e-6), "row scales differ"
        # the kernel multiplies by 1/scale, the reference divides: values on a rounding boundary may land on the adjacent
        # e4m3 code.  Allow <= 0.5 % such elements, each exactly one code apart.
        mism = (q != q_ref).float().mean().item()
        step = (q.view(torch.uint8).int() - q_ref.view(torch.uint8).int()).abs().max().item()
        assert mism < 5e-3 and step <= 1, f"quantised bytes differ on {mism:.2%} of the elements, max code distance {step}"
        qt, st = C.fp8_quantize_cols_transposed(x)
        fp8.use_native_kernels(False)
        qt_ref, st_ref = fp8.quantize_cols_transposed(x)
        fp8.use_native_kernels(True)
        assert torch.allclose(st, st_ref, rtol=1e-6) and (qt != qt_ref).float().mean().item() < 5e-3
        wq, ws = C.fp8_quantize_rows(w)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        C.gemm_fp8(q, s, wq, ws, out, False)
        ref_recipe = (q.view(torch.float8_e4m3fn).float() @ wq.view(torch.float8_e4m3fn).float().t()) * s[:, None] * ws[None, :]
        exact = x.float() @ w.float().t()
        print(f"{M}x{N}x{K}: kernel vs recipe reference {rel(out, ref_recipe):.2e}, vs fp32 product {rel(out, exact):.2e}", flush=True)
        assert rel(out, ref_recipe) < 5e-3 and rel(out, exact) < 0.06
        acc = torch.ones(M, N, device="cuda")
        C.gemm_fp8(q, s, wq, ws, acc, True)
        assert rel(acc - 1.0, ref_recipe) < 5e-3


def timing():
    C = ops.C()
    for (M, N, K) in [(8192, 6144, 4096), (8192, 28672, 4096), (8192, 4096, 14336)]:
        x = torch.randn(M, K, device="cuda").bfloat16()
        w = torch.randn(N, K, device="cuda").bfloat16()
        xq, xs = C.fp8_quantize_rows(x)
        wq, ws = C.fp8_quantize_rows(w)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

        def t(fn, n=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        t8 = t(lambda: C.gemm_fp8(xq, xs, wq, ws, out, False))
        t16 = t(lambda: C.gemm(x, w, out, True, True, False))
        tq = t(lambda: C.fp8_quantize_rows(x))
        fl = 2.0 * M * N * K
        print(f"{M}x{N}x{K}: fp8 {t8:.3f} ms ({fl / t8 / 1e9:.0f} TFLOP/s) | bf16 {t16:.3f} ms ({fl / t16 / 1e9:.0f} TFLOP/s) | "
              f"quantise X {tq:.3f} ms", flush=True)


if __name__ == "__main__":
    numerics()
    timing()
    print("FP8_OK")
