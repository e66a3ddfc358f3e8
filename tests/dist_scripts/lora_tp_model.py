"""LoRA adapters on every parallel linear under tensor (+ sequence) parallelism against the single-process model: loss and
the gradient of every adapter tensor (the base model is frozen)."""
from common import *
from neuronx_distributed_training_b200.models import lora
from neuronx_distributed_training_b200.models.llama import LlamaConfig, LlamaForCausalLM
from neuronx_distributed_training_b200.parallel.layers import shard_master_weight

TP = int(os.environ.get("TP", "2"))
rank, world = init(tp=TP)
g = torch.Generator().manual_seed(5)
ids = torch.randint(0, 64, (2, 8 * TP), generator=g)
base = dict(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
            num_key_value_heads=2, torch_dtype="float32", use_flash_attention=False, use_cpu_initialization=True, fuse_qkv=True)
lcfg = lora.LoraConfig(lora_rank=4, lora_alpha=8, lora_dropout=0.0, target_modules=["o_proj", "gate_up_proj", "down_proj"])   # (the fused QKV weight has unequal q / k / v blocks: no plain re-sharding rule for the test to copy B)

# single-process reference with random (non-zero) B matrices
ps.destroy_model_parallel(); ps.initialize_model_parallel(1, 1, 1)
torch.manual_seed(11)
ref = LlamaForCausalLM(LlamaConfig(**base))
lora.inject_lora(ref, lcfg)
gen = torch.Generator().manual_seed(9)
with torch.no_grad():
    for n, p in ref.named_parameters():
        if ".adapter.lora_B" in n:
            p.copy_(torch.randn(p.shape, generator=gen) * 0.05)
l1, _ = ref(input_ids=ids, labels=ids)
l1.backward()
ref_params = dict(ref.named_parameters())

for sp in (False, True):
    ps.destroy_model_parallel(); ps.initialize_model_parallel(TP, 1, 1)
    tp, tpr = ps.get_tensor_model_parallel_size(), ps.get_tensor_model_parallel_rank()
    torch.manual_seed(11)
    m = LlamaForCausalLM(LlamaConfig(**base, sequence_parallel_enabled=sp))
    lora.inject_lora(m, lcfg)
    params = dict(m.named_parameters())
    assert set(params) == set(ref_params)
    with torch.no_grad():
        for n, p in params.items():
            if ".adapter." not in n:
                continue                                   # base weights: identical by seeded full-matrix initialisation
            full = ref_params[n].detach()
            p.copy_(shard_master_weight(full, p.partition_dim, tp, tpr, getattr(p, "partition_stride", 1))
                    if getattr(p, "tensor_model_parallel", False) else full)
    loss, _ = m(input_ids=ids, labels=ids)
    loss.backward()
    check(f"lora loss sp={sp}", loss.detach(), l1.detach(), atol=1e-5)
    n_checked = 0
    for n, p in params.items():
        if not p.requires_grad:
            assert ".adapter." not in n
            continue
        gr = p.grad.clone()
        if getattr(p, "sequence_parallel_enabled", False):
            dist.all_reduce(gr, group=ps.get_tensor_model_parallel_group())
        want = ref_params[n].grad
        if getattr(p, "tensor_model_parallel", False):
            want = shard_master_weight(want, p.partition_dim, tp, tpr, getattr(p, "partition_stride", 1))
        check(f"lora grad {n} sp={sp}", gr, want, atol=3e-5)
        n_checked += 1
    assert n_checked == 2 * 3 * 2, n_checked              # A and B of three linears in two layers
finish()
