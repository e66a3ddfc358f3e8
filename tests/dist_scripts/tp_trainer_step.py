"""Two optimizer steps of the HF-Llama task module with TP=2 + SP + ZeRO-1 machinery (SP-gradient sync, global grad-norm
with TP-duplicate filtering, clipping, AdamW) against the single-process module: same losses and the same global norm."""
from common import *
from neuronx_distributed_training_b200.config import from_dict, process_config
from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule

TP = int(os.environ.get("TP", "2"))
rank, world = init(tp=TP)
DP = world // TP                                  # > 1: ZeRO-1 sharding of the optimizer state over the data-parallel group
dpr = ps.get_data_parallel_rank()


def cfg(tp, dp=1):
    return process_config(from_dict({
        "name": "hf_llama", "model_source": "hf", "seed": 5,
        "trainer": {"devices": tp * dp, "num_nodes": 1, "max_steps": 10, "gradient_clip_val": 0.5},
        "distributed_strategy": {"tensor_model_parallel_size": tp, "zero1": True, "sequence_parallel": tp > 1,
                                 "kv_replicator": max(1, tp // 2)},      # TP 4 with 2 KV heads: every KV head lives on two ranks
        "data": {"micro_batch_size": 1, "global_batch_size": 2 * DP},
        "model": {"model_config": "llama_125m", "encoder_seq_length": 16, "num_layers": 2, "hidden_size": 32, "intermediate_size": 64,
                  "num_attention_heads": 4, "num_kv_heads": 2, "vocab_size": 64, "fusions": {"flash_attention": False},
                  "use_cpu_initialization": True, "force_cpu_initialization": True,
                  "optim": {"name": "adamw", "lr": 5e-2, "weight_decay": 0.01, "betas": [0.9, 0.95]}},
        "precision": {"type": "fp32"},
    }), tp * dp)


g = torch.Generator().manual_seed(1)
batches = [torch.randint(0, 64, (2 * DP, 16), generator=g) for _ in range(2)]
mod = HFLlamaModule(cfg(TP, DP), None); mod.setup(); mod.configure_optimizers(10)
assert mod.optimizer.zero1 == (DP > 1)
ours = []
for ids_all in batches:
    ids = ids_all[dpr * 2:(dpr + 1) * 2]
    loss = mod.training_step({"input_ids": ids, "labels": ids.clone()})
    ours.append((loss.detach().clone(), mod.optimizer.grad_norm.detach().clone(), mod.parameter_norm().detach().clone()))
# single-process reference on every rank (no process groups involved)
ps.destroy_model_parallel()
_saved = dist.is_initialized
dist.is_initialized = lambda: False
try:
    ps.initialize_model_parallel(1, 1, 1)
    ref = HFLlamaModule(cfg(1), None); ref.setup(); ref.configure_optimizers(10)
    for (l, gn, pn), ids in zip(ours, batches):
        lr_ = ref.training_step({"input_ids": ids, "labels": ids.clone()})
        check("loss", l, lr_, atol=2e-5)
        check("global grad norm", gn, ref.optimizer.grad_norm, atol=1e-4)
        check("parameter norm", pn, ref.parameter_norm(), atol=1e-4)
finally:
    dist.is_initialized = _saved
finish()
