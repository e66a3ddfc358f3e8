"""Tiny driver for ncu: a few launches of the flagship GEMM shapes (fwd nt / dgrad nn / wgrad tn-fp32-add)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from neuronx_distributed_training_b200 import ops
C = ops.C()
M, N, K = 8192, 28672, 4096
x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); dx = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
dw = torch.zeros(N, K, device="cuda")
for _ in range(3):
    C.gemm(x, w, y, True, True, False)
    C.gemm(y, w, dx, True, False, False)
    C.gemm(y, x, dw, False, False, True)
torch.cuda.synchronize()
