"""In-kernel timeline (clock64 stamps) of the heaviest CTA of the flash backward kernel at the Llama-3-8B shape."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from neuronx_distributed_training_b200 import ops
C = ops.C()
B, S, Hq, Hk = 1, 8192, 32, 8
q = torch.randn(B, S, Hq, 128, device="cuda").bfloat16(); k = torch.randn(B, S, Hk, 128, device="cuda").bfloat16()
v = torch.randn(B, S, Hk, 128, device="cuda").bfloat16(); do = torch.randn_like(q)
sc = 1 / math.sqrt(128)
o, lse = C.flash_attn_fwd(q, k, v, True, 0, sc)
for _ in range(2): C.flash_attn_bwd(do, q, k, v, o, lse, True, 0, sc)
dbg = torch.zeros(4 * 256 * 4, dtype=torch.int64, device="cuda")
C.flash_bwd2_set_debug(dbg)
C.flash_attn_bwd(do, q, k, v, o, lse, True, 0, sc); torch.cuda.synchronize()
C.flash_bwd2_set_debug(torch.empty(0, dtype=torch.int64, device="cuda"))
d = dbg.cpu().view(4, 256, 4)
t0 = d[0, 0, 0].item()
print("step | compute: S seen, P pub, dP seen, dS pub | writer: dQ seen, regs, half0 issued, half1 issued | A: wait P, P seen, dV issued, dP(t+1) issued | B: dS seen, dK issued, dQ free, dQ issued")
for t in list(range(0, 4)) + list(range(100, 106)) + list(range(250, 256)):
    print(t, *[[d[r, t, c].item() - t0 for c in range(4)] for r in range(4)])
a, b = 20, 240
per = (d[0, b, 0].item() - d[0, a, 0].item()) / (b - a)
def avg(r, c1, r2, c2, shift=0):
    return sum(d[r2, t + shift, c2].item() - d[r, t, c1].item() for t in range(a, b)) / (b - a)
print(f"period per 64-query step {per:.0f} cycles (tensor work 1280)")
print(f"compute: S seen->P pub {avg(0,0,0,1):.0f}, P pub->dP seen {avg(0,1,0,2):.0f}, dP seen->dS pub {avg(0,2,0,3):.0f}, dS pub->next S seen {avg(0,3,0,0,1):.0f}")
print(f"writer: dQ seen->regs {avg(1,0,1,1):.0f}, regs->half0 {avg(1,1,1,2):.0f}, half0->half1 {avg(1,2,1,3):.0f}, half1->next dQ seen {avg(1,3,1,0,1):.0f}")
print(f"issuer A: wait P {avg(2,0,2,1):.0f}, dV issue {avg(2,1,2,2):.0f}, wait dS + dP issue {avg(2,2,2,3):.0f}, S(t+2) issue {avg(2,3,2,0,1):.0f}")
print(f"issuer B: dK issue {avg(3,0,3,1):.0f}, wait dQ free {avg(3,1,3,2):.0f}, dQ issue {avg(3,2,3,3):.0f}, wait dS(t+1) {avg(3,3,3,0,1):.0f}")
