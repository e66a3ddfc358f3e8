"""Whole Llama under tensor parallelism (with and without sequence parallelism, fused and unfused LM-head loss) against
the single-process model: same loss and the same gradient for EVERY parameter (replicated ones summed over the ranks
that saw different tokens, sharded ones compared with the matching slice of the reference gradient)."""
from common import *
from neuronx_distributed_training_b200.models.llama import LlamaConfig, LlamaForCausalLM
from neuronx_distributed_training_b200.parallel.layers import shard_master_weight

TP = int(os.environ.get("TP", "2"))
rank, world = init(tp=TP)
tp, tpr = ps.get_tensor_model_parallel_size(), ps.get_tensor_model_parallel_rank()
g = torch.Generator().manual_seed(5)
ids = torch.randint(0, 64, (2, 8 * TP), generator=g)
base = dict(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
            num_key_value_heads=max(2, TP), torch_dtype="float32", use_flash_attention=False, use_cpu_initialization=True, fuse_qkv=False)
runs = {}
for sp in (False, True):
    for fused_ce in (True, False):
        torch.manual_seed(11)
        m = LlamaForCausalLM(LlamaConfig(**base, sequence_parallel_enabled=sp, fused_linear_ce=fused_ce))
        loss, _ = m(input_ids=ids, labels=ids)
        loss.backward()
        grads = {}
        for n, p in m.named_parameters():
            gr = p.grad.clone()
            if getattr(p, "sequence_parallel_enabled", False):
                dist.all_reduce(gr)                      # each TP rank saw S/TP of the tokens (the optimizer does this sum)
            grads[n] = (gr, bool(getattr(p, "tensor_model_parallel", False)), getattr(p, "partition_dim", -1),
                        getattr(p, "partition_stride", 1))
        runs[(sp, fused_ce)] = (loss.detach().clone(), grads)
# DPO path: sequence log-probabilities of chosen || rejected through the vocab-parallel logits, and a DPO-shaped loss on them
from neuronx_distributed_training_b200.trainer.dpo import sequence_logps, dpo_loss
lab = ids.clone(); lab[:, :3] = -100
cat = {"input_ids": ids, "attention_mask": torch.ones_like(ids), "labels": lab}
def dpo_like(model):
    c, r = sequence_logps(model, cat)
    loss, _, _ = dpo_loss(c, r, torch.zeros_like(c) - 1.0, torch.zeros_like(r) - 2.0, beta=0.3)
    return c.detach().clone(), r.detach().clone(), loss
dpo_runs = {}
for sp in (False, True):
    torch.manual_seed(11)
    m = LlamaForCausalLM(LlamaConfig(**base, sequence_parallel_enabled=sp))
    c, r, loss = dpo_like(m)
    loss.backward()
    gn = m.model.norm.weight.grad.clone()
    if sp:
        dist.all_reduce(gn)
    dpo_runs[sp] = (c, r, loss.detach().clone(), gn)
ps.destroy_model_parallel(); ps.initialize_model_parallel(1, 1, 1)
torch.manual_seed(11)
ref = LlamaForCausalLM(LlamaConfig(**base))
c1, r1, ld = dpo_like(ref)
ld.backward()
for sp, (c, r, loss, gn) in dpo_runs.items():
    check(f"dpo chosen logps sp={sp}", c, c1, atol=2e-5)
    check(f"dpo rejected logps sp={sp}", r, r1, atol=2e-5)
    check(f"dpo loss sp={sp}", loss, ld.detach(), atol=1e-5)
    check(f"dpo final-norm grad sp={sp}", gn, ref.model.norm.weight.grad, atol=2e-5)
ref.zero_grad()
torch.manual_seed(11)
ref = LlamaForCausalLM(LlamaConfig(**base))
l1, _ = ref(input_ids=ids, labels=ids)
l1.backward()
rg = {n: p.grad for n, p in ref.named_parameters()}
# tied input / output embeddings: one vocabulary-sharded tensor receives the gradient of both uses
ps.destroy_model_parallel(); ps.initialize_model_parallel(TP, 1, 1)
torch.manual_seed(11)
mt = LlamaForCausalLM(LlamaConfig(**base, sequence_parallel_enabled=True, tie_word_embeddings=True))
lt, _ = mt(input_ids=ids, labels=ids)
lt.backward()
assert mt.lm_head.weight is mt.model.embed_tokens.weight
gt = mt.model.embed_tokens.weight.grad.clone()
ps.destroy_model_parallel(); ps.initialize_model_parallel(1, 1, 1)
torch.manual_seed(11)
rt = LlamaForCausalLM(LlamaConfig(**base, tie_word_embeddings=True))
lrt, _ = rt(input_ids=ids, labels=ids)
lrt.backward()
check("tied loss", lt.detach(), lrt.detach(), atol=1e-5)
check("tied embedding grad", gt, shard_master_weight(rt.model.embed_tokens.weight.grad, 0, tp, tpr), atol=2e-5)
for key, (loss, grads) in runs.items():
    check(f"loss {key}", loss, l1.detach(), atol=1e-5)
    assert set(grads) == set(rg), (sorted(set(grads) ^ set(rg))[:4])
    for n, (gr, is_tp, dim, stride) in grads.items():
        want = shard_master_weight(rg[n], dim, tp, tpr, stride) if is_tp else rg[n]
        check(f"grad {n} sp={key[0]} fused_ce={key[1]}", gr, want, atol=2e-5)
finish()
