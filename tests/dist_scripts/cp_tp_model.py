"""Context parallelism composed with tensor parallelism (TP=2 x CP=2 on 4 ranks): ring and all-to-all attention both
reproduce the single-process loss and the gradients of the replicated (norm) parameters."""
from common import *
from neuronx_distributed_training_b200.models.llama import LlamaConfig, LlamaForCausalLM
from neuronx_distributed_training_b200.data.base import get_batch_on_this_context_parallel_rank

TP, CP = 2, 2
g = torch.Generator().manual_seed(5)
ids = torch.randint(0, 64, (2, 8 * CP), generator=g)
base = dict(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
            num_key_value_heads=2, torch_dtype="float32", use_flash_attention=False, use_cpu_initialization=True)
results = {}
rank, world = init(tp=TP, cp=CP)
for mode, sp in (("ring", False), ("ulysses", False), ("ring", True), ("ulysses", True)):
    cfg = LlamaConfig(**base, use_ring_attention=(mode == "ring"), use_ulysses_attention=(mode == "ulysses"),
                      sequence_parallel_enabled=sp)
    torch.manual_seed(11)
    m = LlamaForCausalLM(cfg)
    batch = get_batch_on_this_context_parallel_rank({"input_ids": ids, "labels": ids.clone()})
    loss, _ = m(input_ids=batch["input_ids"], labels=batch["labels"])
    n_loc = (batch["labels"] != -100).sum().float()
    tot = torch.stack([loss.detach() * n_loc, n_loc])
    dist.all_reduce(tot, group=ps.get_context_model_parallel_group())
    (loss * n_loc / tot[1]).backward()
    gn = m.model.norm.weight.grad.clone()
    dist.all_reduce(gn)                       # sum over CP (different tokens) and TP
    # without SP the TP ranks hold replicas of the norm gradient (divide them out); with SP each holds its tokens' share
    results[(mode, sp)] = (tot[0] / tot[1], gn if sp else gn / TP)
ps.destroy_model_parallel(); ps.initialize_model_parallel(1, 1, 1)
torch.manual_seed(11)
m1 = LlamaForCausalLM(LlamaConfig(**base))
l1, _ = m1(input_ids=ids, labels=ids)
l1.backward()
for mode, (l, gn) in results.items():
    check(f"{mode} loss", l, l1.detach(), atol=1e-5)
    check(f"{mode} final-norm grad", gn, m1.model.norm.weight.grad, atol=2e-5)
finish()
