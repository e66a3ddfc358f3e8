"""Tiny ncu target: two launches each of the kernels that have no committed profile yet (flash forward, both variants via
NXDT_FLASH_SPLIT; flash backward preview; fp8 GEMM; grouped expert GEMM).  Usage: see run_round2_ncu.sh."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from neuronx_distributed_training_b200 import ops
C = ops.C()
which = sys.argv[1] if len(sys.argv) > 1 else "flash_fwd"
torch.manual_seed(0)
if which in ("flash_fwd", "flash_bwd"):
    B, S, Hq, Hk = 1, 8192, 8, 2
    q, k, v = (torch.randn(B, S, h, 128, device="cuda").bfloat16() for h in (Hq, Hk, Hk))
    for _ in range(2):
        o, lse = C.flash_attn_fwd(q, k, v, True, 0, 1 / math.sqrt(128))
    if which == "flash_bwd":
        do = torch.randn_like(o)
        for _ in range(2):
            C.flash_attn_bwd_preview(do, q, k, v, o, lse, True, 0, 1 / math.sqrt(128))
elif which == "fp8":
    os.environ["NXDT_FP8_NATIVE"] = "1"
    M, N, K = 8192, 6144, 4096
    x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
    xq, xs = ops.fp8.quantize_rows(x); wq, ws = ops.fp8.quantize_rows(w)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.fp8.gemm(xq, xs, wq, ws, out=out)
elif which == "grouped":
    E, Kd, Nd = 8, 4096, 3584
    counts = torch.tensor([1100, 900, 1300, 700, 1024, 1000, 1168, 1000], device="cuda", dtype=torch.int32)
    rows = int(counts.sum())
    dest, block_expert, offsets, t_cap = ops.grouped_layout(counts, rows)
    a = torch.randn(t_cap, Kd, device="cuda").bfloat16()
    w = torch.randn(E, Kd, Nd, device="cuda").bfloat16()
    for _ in range(2):
        ops.grouped_gemm(a, w, block_expert, offsets)
torch.cuda.synchronize()
print("PROF_TARGET_DONE", which)
