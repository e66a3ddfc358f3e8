"""flash attention forward: correctness + speed vs torch SDPA (fwd) on the Llama-3-8B shape."""
import sys, os, math, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from neuronx_distributed_training_b200 import ops
C = ops.C()
torch.manual_seed(0)
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm()).item()
for (B, S, Hq, Hk, w) in [(1, 256, 4, 1, 0), (2, 1024, 8, 2, 0), (1, 1024, 4, 1, 300), (1, 384, 2, 2, 0), (1, 2048, 4, 2, 700), (2, 4096, 2, 1, 0), (1, 2048, 2, 1, 100), (2, 2048, 32, 8, 0)]:   # the last one is large enough for two-tile CTAs
    q = torch.randn(B, S, Hq, 128, device="cuda").bfloat16(); k = torch.randn(B, S, Hk, 128, device="cuda").bfloat16(); v = torch.randn(B, S, Hk, 128, device="cuda").bfloat16()
    o, lse = C.flash_attn_fwd(q, k, v, True, w, 1 / math.sqrt(128)); torch.cuda.synchronize()
    r, lr = ops.attention_ref(q, k, v, causal=True, window=w or None, return_lse=True)
    print(f"fwd B{B} S{S} Hq{Hq} Hk{Hk} w{w}: rel_err={rel(o, r):.3e} lse_err={(lse - lr).abs().max().item():.3e}", flush=True)
# full (non-causal) attention: the off-diagonal blocks of ring attention (added after round 1's GPU budget; validate first)
for (B, S, Hq, Hk) in [(1, 256, 4, 1), (2, 512, 8, 2), (1, 2048, 4, 4)]:
    q = torch.randn(B, S, Hq, 128, device="cuda").bfloat16(); k = torch.randn(B, S, Hk, 128, device="cuda").bfloat16(); v = torch.randn(B, S, Hk, 128, device="cuda").bfloat16()
    o, lse = C.flash_attn_fwd(q, k, v, False, 0, 1 / math.sqrt(128)); torch.cuda.synchronize()
    r, lr = ops.attention_ref(q, k, v, causal=False, return_lse=True)
    print(f"fwd non-causal B{B} S{S} Hq{Hq} Hk{Hk}: rel_err={rel(o, r):.3e} lse_err={(lse - lr).abs().max().item():.3e}", flush=True)
B, S, Hq, Hk = 1, 8192, 32, 8
q = torch.randn(B, S, Hq, 128, device="cuda").bfloat16(); k = torch.randn(B, S, Hk, 128, device="cuda").bfloat16(); v = torch.randn(B, S, Hk, 128, device="cuda").bfloat16()
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
flops = 4 * B * Hq * S * S * 128 / 2
# the per-GPU shape of the same model under TP = 8 (4 query heads, 1 KV head): small grid -> one-tile CTAs
q8, k8, v8 = q[:, :, :4].contiguous(), k[:, :, :1].contiguous(), v[:, :, :1].contiguous()
t8 = timeit(lambda: C.flash_attn_fwd(q8, k8, v8, True, 0, 1 / math.sqrt(128)))
t8s = timeit(lambda: F.scaled_dot_product_attention(q8.transpose(1, 2), k8.transpose(1, 2), v8.transpose(1, 2), is_causal=True, enable_gqa=True))
print(f"fwd 8B shape at TP8 (4/1 heads): native {t8:.3f} ms {flops / 8 / t8 / 1e9:.0f} TFLOP/s | SDPA {t8s:.3f} ms {flops / 8 / t8s / 1e9:.0f} TFLOP/s", flush=True)
t = timeit(lambda: C.flash_attn_fwd(q, k, v, True, 0, 1 / math.sqrt(128)))
qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
ts = timeit(lambda: F.scaled_dot_product_attention(qt, kt, vt, is_causal=True, enable_gqa=True))
print(f"fwd 8B shape: native {t:.3f} ms {flops / t / 1e9:.0f} TFLOP/s | SDPA {ts:.3f} ms {flops / ts / 1e9:.0f} TFLOP/s", flush=True)
qg, kg, vg = (x.clone().requires_grad_(True) for x in (qt, kt, vt))
def fb():
    o = F.scaled_dot_product_attention(qg, kg, vg, is_causal=True, enable_gqa=True); o.backward(torch.ones_like(o))
tb = timeit(fb)
print(f"SDPA fwd+bwd {tb:.3f} ms ({3.5 * flops / tb / 1e9:.0f} TFLOP/s at 3.5x fwd flops)", flush=True)
