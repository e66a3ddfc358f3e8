"""One forward + backward of the native flash attention at the Llama-3-8B shape (for ncu captures)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from neuronx_distributed_training_b200 import ops
C = ops.C()
B, S, Hq, Hk = 1, 8192, 32, 8
q = torch.randn(B, S, Hq, 128, device="cuda").bfloat16(); k = torch.randn(B, S, Hk, 128, device="cuda").bfloat16()
v = torch.randn(B, S, Hk, 128, device="cuda").bfloat16(); do = torch.randn_like(q)
sc = 1 / math.sqrt(128)
for _ in range(2):
    o, lse = C.flash_attn_fwd(q, k, v, True, 0, sc)
    dq, dk, dv = C.flash_attn_bwd(do, q, k, v, o, lse, True, 0, sc)
torch.cuda.synchronize()
print("done")
