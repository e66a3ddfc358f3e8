"""All-to-all (Ulysses) context parallelism == causal attention on the full sequence; also the full-model CP path vs CP=1.
KV heads (2) are fewer than CP=4 in the second parametrisation: exercises the KV-head replication."""
from common import *
from neuronx_distributed_training_b200 import ops
from neuronx_distributed_training_b200.parallel.ulysses_attention import ulysses_attention
from neuronx_distributed_training_b200.models.llama import LlamaConfig, LlamaForCausalLM
from neuronx_distributed_training_b200.data.base import get_batch_on_this_context_parallel_rank

CP = int(os.environ.get("CP", "2"))
rank, world = init(cp=CP)
cpr = ps.get_context_model_parallel_rank()
g = torch.Generator().manual_seed(3)
B, S, H, Hk, D = 2, 8 * CP, 4, 2, 8
q = torch.randn(B, S, H, D, generator=g); k = torch.randn(B, S, Hk, D, generator=g); v = torch.randn(B, S, Hk, D, generator=g)
do = torch.randn(B, S, H, D, generator=g)
qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
ref = ops.attention_ref(qr, kr, vr, causal=True)
ref.backward(do)
c = S // CP
sl = slice(cpr * c, (cpr + 1) * c)
ql, kl, vl = (t[:, sl].clone().requires_grad_(True) for t in (q, k, v))
out = ulysses_attention(ql, kl, vl, lambda a, b, c_: ops.attention_ref(a, b, c_, causal=True))
out.backward(do[:, sl])
check("ulysses out", out, ref[:, sl], atol=1e-5)
check("ulysses dq", ql.grad, qr.grad[:, sl], atol=1e-5)
check("ulysses dk", kl.grad, kr.grad[:, sl], atol=1e-5)
check("ulysses dv", vl.grad, vr.grad[:, sl], atol=1e-5)

# full model: CP loss (mean over CP ranks) == single-device loss, grads (summed over CP) == single-device grads
cfg = LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                  num_key_value_heads=2, torch_dtype="float32", use_flash_attention=False, use_ulysses_attention=True, use_cpu_initialization=True)
torch.manual_seed(11)
m = LlamaForCausalLM(cfg)
ids = torch.randint(0, 64, (2, 8 * CP), generator=g)
batch = get_batch_on_this_context_parallel_rank({"input_ids": ids, "labels": ids.clone()})
loss, _ = m(input_ids=batch["input_ids"], labels=batch["labels"])
# token-weighted combination of the per-rank means reproduces the global mean
n_loc = (batch["labels"] != -100).sum().float()
num = loss.detach() * n_loc
tot = torch.stack([num, n_loc]); dist.all_reduce(tot, group=ps.get_context_model_parallel_group())
(loss * n_loc / tot[1]).backward()
ps.destroy_model_parallel(); ps.initialize_model_parallel(1, 1, 1)     # reference: no CP
cfg1 = LlamaConfig(**{**cfg.to_dict(), "use_ulysses_attention": False})
torch.manual_seed(11)
m1 = LlamaForCausalLM(cfg1)
m1.load_state_dict(m.state_dict())
l1, _ = m1(input_ids=ids, labels=ids)
l1.backward()
check("cp loss", tot[0] / tot[1], l1.detach(), atol=1e-5)
for (n, p), (_, p1) in zip(m.named_parameters(), m1.named_parameters()):
    gsum = p.grad.clone(); dist.all_reduce(gsum)
    check(f"cp grad {n}", gsum, p1.grad, atol=2e-5)
finish()
