"""Megatron-GPT (Llama style, untied embeddings) through the 1F1B pipeline (PP=2): three optimizer steps reproduce the
single-process losses (weights copied by name: layer numbering is global)."""
from common import *
from neuronx_distributed_training_b200.config import from_dict, process_config
from neuronx_distributed_training_b200.trainer.megatron_module import MegatronGPTModel as Module

rank, world = init(pp=2)


def cfg(pp):
    return process_config(from_dict({
        "name": "megatron_llama", "model_source": "megatron", "seed": 5,
        "trainer": {"devices": pp, "num_nodes": 1, "max_steps": 10, "gradient_clip_val": 1.0},
        "exp_manager": {"create_checkpoint_callback": False, "create_tensorboard_logger": False},
        "distributed_strategy": {"tensor_model_parallel_size": 1, "pipeline_model_parallel_size": pp, "zero1": True},
        "data": {"micro_batch_size": 1, "global_batch_size": 4, "seq_length": 16},
        "model": {"encoder_seq_length": 16, "max_position_embeddings": 16, "num_layers": 4, "hidden_size": 32, "ffn_hidden_size": 64,
                  "num_attention_heads": 4, "vocab_size": 64, "init_method_std": 0.05, "hidden_dropout": 0.0, "attention_dropout": 0.0,
                  "normalization": "rmsnorm", "position_embedding_type": "rope", "activation": "swiglu", "has_bias": False,
                  "share_embeddings_and_output_weights": False, "make_vocab_size_divisible_by": 8, "fusions": {"flash_attention": False},
                  "optim": {"name": "adamw", "lr": 1e-2, "weight_decay": 0.0, "betas": [0.9, 0.95]}},
        "precision": {"type": "fp32"},
    }), pp)


g = torch.Generator().manual_seed(1)
def mk():
    toks = torch.randint(0, 64, (4, 17), generator=g)
    return {"tokens": toks[:, :-1].contiguous(), "labels": toks[:, 1:].contiguous(), "loss_mask": torch.ones(4, 16),
            "position_ids": torch.arange(16).unsqueeze(0).expand(4, 16).contiguous()}
batches = [mk() for _ in range(3)]
# single-process reference first (its initial weights are copied into the pipeline stages by name: global layer numbering)
ps.destroy_model_parallel()
_saved = dist.is_initialized
dist.is_initialized = lambda: False
try:
    ps.initialize_model_parallel(1, 1, 1)
    ref = Module(cfg(1), None); ref.setup(); ref.configure_optimizers(10)
    init_state = {n: p.detach().clone() for n, p in ref.local_named_parameters()}
    theirs = [float(ref.training_step(dict(ids))) for ids in batches]
finally:
    dist.is_initialized = _saved
ps.destroy_model_parallel(); ps.initialize_model_parallel(1, 2, 1)
mod = Module(cfg(2), None); mod.setup()
with torch.no_grad():
    for n, p in mod.local_named_parameters():
        p.copy_(init_state[n])
mod.configure_optimizers(10)
ours = []
for i, ids in enumerate(batches):
    loss = mod.training_step(dict(ids))
    ours.append(float(loss))
if rank == 0:
    print("pp losses", ours, "single", theirs, flush=True)
for a, b in zip(ours, theirs):
    assert abs(a - b) < 2e-5, (ours, theirs)
finish()
