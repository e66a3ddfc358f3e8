"""Expert parallelism: Mixtral with EP=2 over DP=2 trains identically to the EP=1 single-process model on the
concatenated batch (same loss, same parameter update)."""
from common import *
from neuronx_distributed_training_b200.config import from_dict, process_config
from neuronx_distributed_training_b200.trainer.hf_mixtral_module import HFMixtralModule

EP = int(os.environ.get("EP", "2"))
TP = int(os.environ.get("TP", "1"))          # TP > 1: experts are additionally sharded on the FFN dimension (+ sequence parallel)
rank, world = init(tp=TP, ep=EP)
def cfg(ep, dp_world, tp=1):
    return from_dict({
        "name": "hf_mixtral", "model_source": "hf", "seed": 5,
        "trainer": {"devices": dp_world * tp, "num_nodes": 1, "max_steps": 10, "gradient_clip_val": 1.0},
        "distributed_strategy": {"tensor_model_parallel_size": tp, "expert_model_parallel_size": ep, "zero1": True,
                                 "sequence_parallel": tp > 1},
        "data": {"micro_batch_size": 1, "global_batch_size": dp_world},
        "model": {"model_config": "mixtral_8x7b", "encoder_seq_length": 16, "num_layers": 2, "hidden_size": 32, "intermediate_size": 48,
                  "num_attention_heads": 4, "num_kv_heads": 2, "vocab_size": 64, "fusions": {"flash_attention": False},
                  "force_cpu_initialization": True,
                  "moe": {"num_experts": 4, "top_k": 2, "dropless": True, "router_aux_loss_coef": 0.02, "glu_mlp": True},
                  "optim": {"name": "adamw", "lr": 1e-2, "weight_decay": 0.0, "betas": [0.9, 0.95]}},
        "precision": {"type": "fp32"},
    })
dp_world = world // TP
c = process_config(cfg(EP, dp_world, TP), world)
mod = HFMixtralModule(c, None); mod.setup(); mod.configure_optimizers(10)
g = torch.Generator().manual_seed(1)
ids_all = torch.randint(0, 64, (dp_world, 16), generator=g)
dpr = ps.get_data_parallel_rank()
loss = mod.training_step({"input_ids": ids_all[dpr:dpr + 1], "labels": ids_all[dpr:dpr + 1].clone()})
gn = mod.optimizer.grad_norm.clone()
# reference: same seed, EP=1, single process sees the whole batch as `world` microbatches
ps.destroy_model_parallel(); 
import torch.distributed as dist
# build the reference on every rank without process groups (pure local computation)
_saved = dist.is_initialized
dist.is_initialized = lambda: False
try:
    ps.initialize_model_parallel(1, 1, 1, 1)
    c1 = process_config(cfg(1, 1), 1)
    c1.data["global_batch_size"] = dp_world
    ref = HFMixtralModule(c1, None); ref.setup(); ref.configure_optimizers(10)
    lref = ref.training_step({"input_ids": ids_all, "labels": ids_all.clone()})
finally:
    dist.is_initialized = _saved
check("ep loss", loss, lref, atol=1e-5)
check("ep grad norm", gn, ref.optimizer.grad_norm, atol=1e-4)
ref_params = dict(ref.local_named_parameters())
El = 4 // EP
epr = (rank // TP) % EP                       # TP varies fastest in the rank grid
from neuronx_distributed_training_b200.parallel.layers import shard_master_weight
tp_, tpr_ = TP, rank % TP
for n, p in mod.local_named_parameters():
    r = ref_params[n]
    if getattr(p, "expert_parallel", False):
        r = r[epr * El:(epr + 1) * El]
    if TP > 1 and getattr(p, "tensor_model_parallel", False):
        if "qkv_proj" in n:
            continue                              # fused q/k/v blocks of unequal size: no plain re-sharding rule (covered by tp_model.py)
        r = shard_master_weight(r, p.partition_dim, tp_, tpr_, getattr(p, "partition_stride", 1))
    check(f"ep param {n}", p, r, atol=2e-5)
finish()
