import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from neuronx_distributed_training_b200 import ops
C = ops.C()
torch.manual_seed(0)
for (B, S, Hq, Hk, w) in [(1, 2048, 4, 2, 700), (1, 2048, 2, 1, 100), (1, 1024, 4, 1, 300), (1, 2048, 1, 1, 700)]:
    q = torch.randn(B, S, Hq, 128, device="cuda").bfloat16(); k = torch.randn(B, S, Hk, 128, device="cuda").bfloat16(); v = torch.randn(B, S, Hk, 128, device="cuda").bfloat16()
    for rep in range(3):
        o, lse = C.flash_attn_fwd(q, k, v, True, w, 1 / math.sqrt(128)); torch.cuda.synchronize()
        bad = torch.isnan(o.float()).any(-1)          # [B,S,H]
        rows = bad.any(-1)[0].nonzero().flatten()
        r, _ = ops.attention_ref(q, k, v, causal=True, window=w, return_lse=True)
        err = ((o.float() - r.float()).norm(dim=-1) / r.float().norm(dim=-1).clamp_min(1e-6))[0]      # [S,H]
        big = (err > 0.05).any(-1).nonzero().flatten()
        print(f"S{S} Hq{Hq} w{w} rep{rep}: nan rows {rows.numel()} [{rows[:6].tolist()} .. {rows[-3:].tolist() if rows.numel() else []}] heads {bad.any(1)[0].tolist()}"
              f" | rows with err>5%: {big.numel()} {big[:8].tolist()}", flush=True)
