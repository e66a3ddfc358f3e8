"""Megatron-GPT under tensor + sequence parallelism against the single-process model: loss and every parameter gradient
(GPT-2 style: learned positions, LayerNorm, biases, GeLU, tied embeddings; Llama style: RoPE, RMSNorm, SwiGLU, untied)."""
from common import *
from neuronx_distributed_training_b200.config import from_dict, process_config
from neuronx_distributed_training_b200.parallel.layers import shard_master_weight
from neuronx_distributed_training_b200.trainer.megatron_module import MegatronGPTModel

TP = int(os.environ.get("TP", "2"))
rank, world = init(tp=TP)


def cfg(tp, model_kw):
    model = {"encoder_seq_length": 16, "max_position_embeddings": 16, "num_layers": 2, "hidden_size": 32, "ffn_hidden_size": 64,
             "num_attention_heads": 4, "vocab_size": 64, "init_method_std": 0.05, "hidden_dropout": 0.0, "attention_dropout": 0.0,
             "normalization": "layernorm", "position_embedding_type": "learned_absolute", "activation": "gelu", "has_bias": True,
             "share_embeddings_and_output_weights": True, "make_vocab_size_divisible_by": 8, "fusions": {"flash_attention": False},
             "optim": {"name": "adamw", "lr": 1e-2, "weight_decay": 0.0, "betas": [0.9, 0.95]}}
    model.update(model_kw)
    d = {"name": "megatron_gpt", "model_source": "megatron", "seed": 5,
         "trainer": {"devices": tp, "num_nodes": 1, "max_steps": 5},
         "exp_manager": {"create_checkpoint_callback": False, "create_tensorboard_logger": False},
         "distributed_strategy": {"tensor_model_parallel_size": tp, "zero1": True, "sequence_parallel": tp > 1},
         "data": {"micro_batch_size": 2, "global_batch_size": 2, "seq_length": 16},
         "model": model, "precision": {"type": "fp32"}}
    return process_config(from_dict(d), tp)


g = torch.Generator().manual_seed(3)
toks = torch.randint(0, 64, (2, 17), generator=g)
batch = {"tokens": toks[:, :-1].contiguous(), "labels": toks[:, 1:].contiguous(), "loss_mask": torch.ones(2, 16),
         "position_ids": torch.arange(16).unsqueeze(0).expand(2, 16).contiguous()}
variants = {"gpt2": {},
            "llama": {"normalization": "rmsnorm", "position_embedding_type": "rope", "activation": "swiglu", "has_bias": False,
                      "share_embeddings_and_output_weights": False}}
for vname, kw in variants.items():
    # single-process reference (every rank builds the same one)
    ps.destroy_model_parallel(); ps.initialize_model_parallel(1, 1, 1)
    torch.manual_seed(21)
    ref = MegatronGPTModel(cfg(1, kw), None); ref.setup()
    l1, _ = ref.model_fwd_calc_loss(dict(batch))
    l1.backward()
    ref_params = dict(ref.model.named_parameters())
    # tensor + sequence parallel model with the reference's weights
    ps.destroy_model_parallel(); ps.initialize_model_parallel(TP, 1, 1)
    tp, tpr = ps.get_tensor_model_parallel_size(), ps.get_tensor_model_parallel_rank()
    torch.manual_seed(21)
    mod = MegatronGPTModel(cfg(TP, kw), None); mod.setup()
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
        gr = p.grad.clone()
        if getattr(p, "sequence_parallel_enabled", False):
            dist.all_reduce(gr, group=ps.get_tensor_model_parallel_group())
        want = ref_params[n].grad
        if getattr(p, "tensor_model_parallel", False):
            want = shard_master_weight(want, p.partition_dim, tp, tpr, getattr(p, "partition_stride", 1))
        check(f"{vname} grad {n}", gr, want, atol=3e-5)
finish()
