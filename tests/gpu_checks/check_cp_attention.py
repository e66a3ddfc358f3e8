"""Context-parallel attention on GPUs (torchrun --nproc-per-node CP): gathered-KV native kernel (forward + backward) on
every rank's sequence chunk against the single-GPU native flash attention and the fp32 oracle on the full sequence."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from neuronx_distributed_training_b200 import ops
from neuronx_distributed_training_b200.parallel import state as ps
from neuronx_distributed_training_b200.parallel.ring_attention import ring_attention, _native_cp_ok

ps.init_distributed("nccl")
rank, cp = dist.get_rank(), dist.get_world_size()
ps.initialize_model_parallel(1, 1, cp)
dev = torch.device("cuda", torch.cuda.current_device())
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-9)).item()
ok = True
for (B, S, Hq, Hk) in [(1, 1024 * cp, 8, 2), (2, 512 * cp, 4, 4)]:
    g = torch.Generator(device="cpu").manual_seed(7)
    q = torch.randn(B, S, Hq, 128, generator=g).to(dev).bfloat16()
    k = torch.randn(B, S, Hk, 128, generator=g).to(dev).bfloat16()
    v = torch.randn(B, S, Hk, 128, generator=g).to(dev).bfloat16()
    do = torch.randn(B, S, Hq, 128, generator=g).to(dev).bfloat16()
    Sl = S // cp
    sl = slice(rank * Sl, (rank + 1) * Sl)
    ql, kl, vl = (t[:, sl].contiguous().requires_grad_(True) for t in (q, k, v))
    assert _native_cp_ok(ql, kl)
    n0 = ops.native_launch_count()
    out = ring_attention(ql, kl, vl)
    out.backward(do[:, sl].contiguous())
    assert ops.native_launch_count() - n0 == 4
    qf, kf, vf = (t.clone().requires_grad_(True) for t in (q, k, v))
    ref = ops.attention_ref(qf, kf, vf, causal=True) if S <= 2048 else ops.flash_attention(qf, kf, vf, causal=True)
    ref.backward(do)
    errs = {"out": rel(out, ref[:, sl]), "dq": rel(ql.grad, qf.grad[:, sl]), "dk": rel(kl.grad, kf.grad[:, sl]), "dv": rel(vl.grad, vf.grad[:, sl])}
    print(f"[rank {rank}] CP={cp} B{B} S{S} Hq{Hq} Hk{Hk}: " + " ".join(f"{n} {e:.2e}" for n, e in errs.items()), flush=True)
    ok = ok and all(e < 1.5e-2 for e in errs.values())
# timing at a long-context shape: 32k tokens, 8 q heads / 2 kv heads per rank (Llama-3-70B at TP 8), whole sequence over CP ranks
B, S, Hq, Hk = 1, 32768, 8, 1
Sl = S // cp
ql = torch.randn(B, Sl, Hq, 128, device=dev).bfloat16().requires_grad_(True)
kl = torch.randn(B, Sl, Hk, 128, device=dev).bfloat16().requires_grad_(True)
vl = torch.randn(B, Sl, Hk, 128, device=dev).bfloat16().requires_grad_(True)
dol = torch.randn(B, Sl, Hq, 128, device=dev).bfloat16()
def step():
    ql.grad = kl.grad = vl.grad = None
    ring_attention(ql, kl, vl).backward(dol)
for _ in range(3): step()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 5], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
flops = 3.5 * 4 * B * Hq * S * S * 128 / 2
flag = torch.tensor([0 if ok else 1], device=dev); dist.all_reduce(flag)
if rank == 0:
    print(f"CP={cp} fwd+bwd 32k tokens, 8/1 heads per rank: {t.item():.3f} ms (max over ranks), {flops / t.item() / 1e9 / cp:.0f} TFLOP/s per GPU "
          f"(last rank does {2 * cp / (cp + 1):.2f}x the mean work under contiguous causal chunks)", flush=True)
    print("CP_ATTENTION_OK" if flag.item() == 0 else "CP_ATTENTION_MISMATCH", flush=True)
dist.barrier(); dist.destroy_process_group()
