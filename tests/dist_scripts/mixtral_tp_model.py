"""HF-Mixtral (MoE, TP-sharded experts) under tensor + sequence parallelism against the single-process model: loss (incl.
the router auxiliary loss) and every parameter gradient.  Dropless and capacity-factor routing."""
from common import *
from neuronx_distributed_training_b200.config import from_dict, process_config
from neuronx_distributed_training_b200.parallel.layers import shard_master_weight
from neuronx_distributed_training_b200.trainer.hf_mixtral_module import HFMixtralModule

TP = int(os.environ.get("TP", "2"))
rank, world = init(tp=TP)


def cfg(tp, moe_kw, sp=True):
    return process_config(from_dict({
        "name": "hf_mixtral", "model_source": "hf", "seed": 5,
        "trainer": {"devices": tp, "num_nodes": 1, "max_steps": 10},
        "distributed_strategy": {"tensor_model_parallel_size": tp, "zero1": True, "sequence_parallel": tp > 1 and sp},
        "data": {"micro_batch_size": 2, "global_batch_size": 2},
        "model": {"model_config": "mixtral_8x7b", "encoder_seq_length": 16, "num_layers": 2, "hidden_size": 32, "intermediate_size": 48,
                  "num_attention_heads": 4, "num_kv_heads": 2, "vocab_size": 64, "fusions": {"flash_attention": False},
                  "force_cpu_initialization": True, "fuse_qkv": False,
                  "moe": {"num_experts": 4, "top_k": 2, "router_aux_loss_coef": 0.02, "glu_mlp": True, **moe_kw},
                  "optim": {"name": "adamw", "lr": 1e-2, "weight_decay": 0.0, "betas": [0.9, 0.95]}},
        "precision": {"type": "fp32"},
    }), tp)


g = torch.Generator().manual_seed(1)
ids = torch.randint(0, 64, (2, 16), generator=g)
batch = {"input_ids": ids, "labels": ids.clone()}
for vname, moe_kw, sp in (("dropless+sp", {"dropless": True}, True), ("capacity+sp", {"dropless": False, "capacity_factor": 4.0}, True),
                          ("dropless", {"dropless": True}, False)):
    ps.destroy_model_parallel(); ps.initialize_model_parallel(1, 1, 1)
    torch.manual_seed(21)
    ref = HFMixtralModule(cfg(1, moe_kw), None); ref.setup()
    l1, _ = ref.model_fwd_calc_loss(dict(batch))
    l1.backward()
    ref_params = dict(ref.model.named_parameters())
    ps.destroy_model_parallel(); ps.initialize_model_parallel(TP, 1, 1)
    tp, tpr = ps.get_tensor_model_parallel_size(), ps.get_tensor_model_parallel_rank()
    torch.manual_seed(21)
    mod = HFMixtralModule(cfg(TP, moe_kw, sp), None); mod.setup()
    params = dict(mod.model.named_parameters())
    assert set(params) == set(ref_params), sorted(set(params) ^ set(ref_params))[:4]
    with torch.no_grad():
        for n, p in params.items():
            full = ref_params[n].detach()
            p.copy_(shard_master_weight(full, p.partition_dim, tp, tpr, getattr(p, "partition_stride", 1))
                    if getattr(p, "tensor_model_parallel", False) else full)
    loss, _ = mod.model_fwd_calc_loss(dict(batch))
    loss.backward()
    check(f"{vname} loss", loss.detach(), l1.detach(), atol=1e-5)
    for n, p in params.items():
        assert p.grad is not None, n
        gr = p.grad.clone()
        if getattr(p, "sequence_parallel_enabled", False):
            dist.all_reduce(gr, group=ps.get_tensor_model_parallel_group())
        want = ref_params[n].grad
        if getattr(p, "tensor_model_parallel", False):
            want = shard_master_weight(want, p.partition_dim, tp, tpr, getattr(p, "partition_stride", 1))
        check(f"{vname} grad {n}", gr, want, atol=3e-5)
finish()
