"""Small-shape launches of every tcgen05 / TMA kernel for `compute-sanitizer` (memcheck / racecheck / synccheck): a few CTAs
each, so that the instrumented run finishes in seconds.  tests/gpu_checks/run_r2_sanitize.sh; results in profiles/sanitizer.md."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from neuronx_distributed_training_b200 import ops
C = ops.C()
torch.manual_seed(0)
dev = "cuda"
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm()).item()
# SM-pair GEMM and one-CTA GEMM, all layouts, store + accumulate epilogues, ragged shape
for M, N, K in ((512, 512, 256), (200, 264, 136)):
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    C.gemm(a, b, out, True, True, False); print("gemm nt", M, N, K, rel(out, a.float() @ b.float().t()))
    bt = b.t().contiguous(); C.gemm(a, bt, out, True, False, False); print("gemm nn", rel(out, a.float() @ b.float().t()))
    at = a.t().contiguous(); acc = torch.zeros(M, N, device=dev); C.gemm(at, bt, acc, False, False, True)
    print("gemm tn f32 add", rel(acc, a.float() @ b.float().t()))
    C.gemm_set_2cta(False); C.gemm(a, b, out, True, True, False); C.gemm_set_2cta(True)
    print("gemm nt one-CTA", rel(out, a.float() @ b.float().t()))
# flash attention forward + backward (two-tile and one-tile forward variants, GQA, window)
for (B, S, Hq, Hk, win) in ((1, 512, 4, 2, 0), (1, 256, 2, 2, 100)):
    q = torch.randn(B, S, Hq, 128, device=dev).bfloat16(); k = torch.randn(B, S, Hk, 128, device=dev).bfloat16()
    v = torch.randn(B, S, Hk, 128, device=dev).bfloat16(); do = torch.randn_like(q)
    o, lse = C.flash_attn_fwd(q, k, v, True, win, 1 / math.sqrt(128))
    dq, dk, dv = C.flash_attn_bwd(do, q, k, v, o, lse, True, win, 1 / math.sqrt(128))
    print("flash", B, S, Hq, Hk, win, bool(torch.isfinite(o.float()).all() and torch.isfinite(dq.float()).all()))
# MX fp8 GEMM + quantisers
a = torch.randn(256, 256, device=dev).bfloat16(); w = torch.randn(512, 256, device=dev).bfloat16()
aq, asf = C.mx_quantize_rows(a); wq, wsf = C.mx_quantize_rows(w)
out = torch.empty(256, 512, device=dev, dtype=torch.bfloat16); C.gemm_mxfp8(aq, asf, wq, wsf, out, False)
print("mxfp8", rel(out, a.float() @ w.float().t()))
# elementwise / optimizer kernels
x = torch.randn(300, 1024, device=dev).bfloat16(); wn = torch.ones(1024, device=dev).bfloat16()
y, rstd = C.rmsnorm_fwd(x, wn, 1e-5); dx, dw = C.rmsnorm_bwd(torch.randn_like(x), x, wn, rstd)
print("rmsnorm", bool(torch.isfinite(dx.float()).all()))
torch.cuda.synchronize()
print("SANITIZE_SMALL_DONE")
