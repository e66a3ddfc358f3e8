"""ONE launch of every headline kernel at its Llama-3-8B shape, in a fixed order, for a single `ncu --set full` capture
(ncu's own cache control makes warm-up unnecessary).  `tests/gpu_checks/run_r2_ncu.sh` is the command; the summaries are in
`profiles/ncu_round2.md`."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from neuronx_distributed_training_b200 import ops
from neuronx_distributed_training_b200.ops import fp8
C = ops.C()
torch.manual_seed(0)
dev = "cuda"
# 1-2. flash attention forward / backward (B1, S8192, 32 / 8 heads, D128, causal)
B, S, Hq, Hk = 1, 8192, 32, 8
q = torch.randn(B, S, Hq, 128, device=dev).bfloat16(); k = torch.randn(B, S, Hk, 128, device=dev).bfloat16()
v = torch.randn(B, S, Hk, 128, device=dev).bfloat16(); do = torch.randn_like(q)
sc = 1 / math.sqrt(128)
o, lse = C.flash_attn_fwd(q, k, v, True, 0, sc)
dq, dk, dv = C.flash_attn_bwd(do, q, k, v, o, lse, True, 0, sc)
# 3-4. bf16 GEMM, one CTA per tile and SM pair per tile (gate_up forward shape)
M, N, K = 8192, 28672, 4096
a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16()
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
C.gemm(a, w, out, True, True, False)
if hasattr(C, "gemm_2cta"):
    C.gemm_2cta(a, w, out, True, True, False)
# 5-7. MX fp8: the two quantisers + the block-scaled GEMM
aq, asf = C.mx_quantize_rows(a); wq, wsf = C.mx_quantize_rows(w)
C.gemm_mxfp8(aq, asf, wq, wsf, out, False)
# 8. wgrad (MN-major operands, fp32 reduce-add epilogue)
dy = torch.randn(M, 4096, device=dev).bfloat16(); x = torch.randn(M, 4096, device=dev).bfloat16()
mg = torch.zeros(4096, 4096, device=dev)
C.gemm(dy, x, mg, False, False, True)
torch.cuda.synchronize()
print("done")
