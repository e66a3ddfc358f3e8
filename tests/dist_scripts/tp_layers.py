"""TP(+SP) layers vs a single-process dense reference: forward, input grads, weight grads."""
from common import *
import torch.nn.functional as F
from neuronx_distributed_training_b200.parallel.layers import (ColumnParallelLinear, RowParallelLinear,
    GQAQKVColumnParallelLinear, ParallelEmbedding, shard_master_weight)
from neuronx_distributed_training_b200.parallel import mappings as mp
from neuronx_distributed_training_b200.parallel.loss import (parallel_cross_entropy, from_parallel_logits_to_logprobs,
                                                              fused_linear_cross_entropy)

rank, world = init(tp=int(os.environ.get("TP", "2")))
tp, tpr = ps.get_tensor_model_parallel_size(), ps.get_tensor_model_parallel_rank()
S, B, H, I = 8 * tp, 2, 16, 24 * tp
g = torch.Generator().manual_seed(7)
x_full = torch.randn(S, B, H, generator=g)
w1 = torch.randn(I, H, generator=g) * 0.1       # column weight (stride 2: gate||up)
w2 = torch.randn(H, I // 2, generator=g) * 0.1  # row weight
gy = torch.randn(S, B, H, generator=g)

from neuronx_distributed_training_b200.parallel import layers as _layers
for sp, keep in ((False, False), (True, False), (True, True)):
    _layers.set_keep_gathered_input(keep)       # keep=True: the gathered input is saved for wgrad (no second all-gather)
    col = ColumnParallelLinear(H, I, bias=False, stride=2, sequence_parallel_enabled=sp)
    row = RowParallelLinear(I // 2, H, bias=False, sequence_parallel_enabled=sp)
    with torch.no_grad():
        col.weight.copy_(shard_master_weight(w1, 0, tp, tpr, stride=2))
        row.weight.copy_(shard_master_weight(w2, 1, tp, tpr))
    x_in = (x_full.chunk(tp, 0)[tpr] if sp else x_full).clone().requires_grad_(True)
    h = col(x_in)
    gate, up = h.chunk(2, -1)
    y = row(F.silu(gate) * up)
    gy_l = gy.chunk(tp, 0)[tpr] if sp else gy
    y.backward(gy_l)
    # dense reference
    xr = x_full.clone().requires_grad_(True)
    w1r, w2r = w1.clone().requires_grad_(True), w2.clone().requires_grad_(True)
    hr = xr @ w1r.t()
    gr, ur = hr.chunk(2, -1)
    yr = (F.silu(gr) * ur) @ w2r.t()
    yr.backward(gy)
    check(f"y sp={sp}", y, yr.chunk(tp, 0)[tpr] if sp else yr)
    check(f"dx sp={sp} keep={keep}", x_in.grad, xr.grad.chunk(tp, 0)[tpr] if sp else xr.grad)
    check(f"dw1 sp={sp} keep={keep}", col.weight.grad, shard_master_weight(w1r.grad, 0, tp, tpr, stride=2))
    check(f"dw2 sp={sp}", row.weight.grad, shard_master_weight(w2r.grad, 1, tp, tpr))
_layers.set_keep_gathered_input(False)

# ---- GQA QKV with KV replication --------------------------------------------------------------------------
nh, nkv, hd = 2 * tp, max(1, tp // 2), 4
mult = 2 if nkv * 1 < tp else 1
for fuse in (True, False):
    torch.manual_seed(5)
    qkv = GQAQKVColumnParallelLinear(H, nh, nkv, hd, kv_size_multiplier=mult, fuse_qkv=fuse, sequence_parallel_enabled=True)
    gq = torch.Generator().manual_seed(11)
    qf, kf, vf = torch.randn(nh * hd, H, generator=gq), torch.randn(nkv * hd, H, generator=gq), torch.randn(nkv * hd, H, generator=gq)
    qkv.load_full_qkv(qf, kf, vf)
    x_in = x_full.chunk(tp, 0)[tpr].clone().requires_grad_(True)
    q, k, v = qkv(x_in)
    qr = x_full @ qf.t()
    kr = x_full @ qkv.replicate_kv(kf).t()
    check(f"gqa q fuse={fuse}", q, qr.chunk(tp, -1)[tpr])
    check(f"gqa k fuse={fuse}", k, kr.chunk(tp, -1)[tpr])
    (q.sum() + k.sum() + v.sum()).backward()
    qkv.sync_kv_replica_grads()

# ---- embedding + vocab-parallel CE -----------------------------------------------------------------------------
V = 32 * tp
ids = torch.randint(0, V, (B, S), generator=g)
emb_w = torch.randn(V, H, generator=g)
for sp in (False, True):
    emb = ParallelEmbedding(V, H, sequence_parallel_enabled=sp)
    with torch.no_grad():
        emb.weight.copy_(shard_master_weight(emb_w, 0, tp, tpr))
    out = emb(ids)
    ref = F.embedding(ids, emb_w)
    if sp:
        check("emb sp", out, ref.transpose(0, 1).chunk(tp, 0)[tpr])
    else:
        check("emb", out, ref)
    out.sum().backward()
    refw = emb_w.clone().requires_grad_(True)
    F.embedding(ids, refw).sum().backward()
    check(f"emb dw sp={sp}", emb.weight.grad, shard_master_weight(refw.grad, 0, tp, tpr))

logits_full = torch.randn(B, S, V, generator=g, requires_grad=True)
labels = torch.randint(0, V, (B, S), generator=g)
labels[0, :3] = -100
ll = logits_full.detach().chunk(tp, -1)[tpr].clone().requires_grad_(True)
loss = parallel_cross_entropy(ll, labels)
ref = F.cross_entropy(logits_full.reshape(-1, V), labels.reshape(-1), reduction="none", ignore_index=-100).view(B, S)
check("pce", loss, ref)
loss.sum().backward(); ref.sum().backward()
check("pce grad", ll.grad, logits_full.grad.chunk(tp, -1)[tpr])
lp = from_parallel_logits_to_logprobs(ll.detach(), labels)
check("logprobs", lp, -ref.detach())
# ... and its backward (the DPO / ORPO losses differentiate through the per-token log-probabilities)
ll2 = logits_full.detach().chunk(tp, -1)[tpr].clone().requires_grad_(True)
wts = torch.randn(B, S, generator=g)
(from_parallel_logits_to_logprobs(ll2, labels) * wts).sum().backward()
lf = logits_full.detach().clone().requires_grad_(True)
lpr = -F.cross_entropy(lf.reshape(-1, V), labels.reshape(-1), reduction="none", ignore_index=-100).view(B, S)
(lpr * wts).sum().backward()
check("logprobs grad", ll2.grad, lf.grad.chunk(tp, -1)[tpr])

# ---- fused linear + CE ------------------------------------------------------------------------------------------
T = 24
hid = torch.randn(T, H, generator=g)
wv = torch.randn(V, H, generator=g) * 0.2
lab = torch.randint(0, V, (T,), generator=g); lab[:2] = -100
h1 = hid.clone().requires_grad_(True)
wl = shard_master_weight(wv, 0, tp, tpr).clone().requires_grad_(True)
# the fused op returns the PARTIAL d(hidden) of its vocabulary slice; the region-entry mapping sums it over TP
ls, n = fused_linear_cross_entropy(mp.copy_to_tensor_model_parallel_region(h1), wl, lab, chunk_tokens=7)
(ls / n).backward()
h2 = hid.clone().requires_grad_(True); w2_ = wv.clone().requires_grad_(True)
r = F.cross_entropy(h2 @ w2_.t(), lab, ignore_index=-100)
r.backward()
check("flce loss", ls / n, r)
check("flce dh", h1.grad, h2.grad)
check("flce dw", wl.grad, shard_master_weight(w2_.grad, 0, tp, tpr))

# ---- fp8 linears: quantise-then-gather (e4m3 bytes + row scales over the wire) == gather-then-quantise, bit for bit ----
from neuronx_distributed_training_b200 import ops
ops.fp8.set_enabled(True)
try:
    col = ColumnParallelLinear(H, I, bias=False, stride=2, sequence_parallel_enabled=True)
    with torch.no_grad():
        col.weight.copy_(shard_master_weight(w1, 0, tp, tpr, stride=2))
    outs = {}
    for mode in ("1", "0"):
        os.environ["NXDT_FP8_GATHER"] = mode
        x_in = x_full.chunk(tp, 0)[tpr].clone().requires_grad_(True)
        y = col(x_in)
        y.backward(torch.ones_like(y))
        outs[mode] = (y.detach().clone(), x_in.grad.clone(), col.weight.grad.clone())
        col.weight.grad = None
    for a, b, name in zip(outs["1"], outs["0"], ("y", "dx", "dw")):
        assert torch.equal(a, b), f"fp8 gather mode changed {name}"
    ref = x_full.reshape(-1, H) @ shard_master_weight(w1, 0, tp, tpr, stride=2).t()
    rel = ((outs["1"][0].reshape(-1, ref.shape[1]) - ref).norm() / ref.norm()).item()
    assert rel < 0.08, rel                      # e4m3 with per-row scales
finally:
    ops.fp8.set_enabled(False)
    os.environ.pop("NXDT_FP8_GATHER", None)
finish()
