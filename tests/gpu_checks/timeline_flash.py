"""In-kernel timeline (clock64 stamps) of the heaviest CTA of the flash forward kernel at the Llama-3-8B shape."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from neuronx_distributed_training_b200 import ops
C = ops.C()
B, S, Hq, Hk = 1, 8192, 32, 8
q = torch.randn(B, S, Hq, 128, device="cuda").bfloat16(); k = torch.randn(B, S, Hk, 128, device="cuda").bfloat16()
v = torch.randn(B, S, Hk, 128, device="cuda").bfloat16()
sc = 1 / math.sqrt(128)
for _ in range(2): C.flash_attn_fwd(q, k, v, True, 0, sc)
dbg = torch.zeros(3 * 128 * 4, dtype=torch.int64, device="cuda")
C.flash_fwd2_set_debug(dbg)
C.flash_attn_fwd(q, k, v, True, 0, sc); torch.cuda.synchronize()
C.flash_fwd2_set_debug(torch.empty(0, dtype=torch.int64, device="cuda"))
d = dbg.cpu().view(3, 128, 4)
t0 = d[0, 0, 0].item()
print("block | WG0: S seen, exp start, exp end, P pub | WG1: same | MMA: P0 seen, PV0 issued, P1 seen, PV1 issued   (cycles since WG0's first S)")
for i in list(range(0, 6)) + list(range(60, 66)) + list(range(120, 128)):
    row = [d[r, i, c].item() - t0 if d[r, i, c].item() else -1 for r in range(3) for c in range(4)]
    print(i, row[0:4], row[4:8], row[8:12])
n = 120
per = (d[0, n, 0].item() - d[0, 4, 0].item()) / (n - 4)
e0 = sum(d[0, i, 2].item() - d[0, i, 1].item() for i in range(4, n)) / (n - 4)
e1 = sum(d[1, i, 2].item() - d[1, i, 1].item() for i in range(4, n)) / (n - 4)
l0 = sum(d[0, i, 1].item() - d[0, i, 0].item() for i in range(4, n)) / (n - 4)
l1 = sum(d[1, i, 1].item() - d[1, i, 0].item() for i in range(4, n)) / (n - 4)
w0 = sum(d[0, i + 1, 0].item() - d[0, i, 3].item() for i in range(4, n)) / (n - 4)
w1 = sum(d[1, i + 1, 0].item() - d[1, i, 3].item() for i in range(4, n)) / (n - 4)
print(f"period per 64-key block {per:.0f} cycles (tensor work 1024); exp phase WG0 {e0:.0f} WG1 {e1:.0f}; S-seen->exp-start WG0 {l0:.0f} WG1 {l1:.0f}; "
      f"P-published->next-S-seen WG0 {w0:.0f} WG1 {w1:.0f}")
