"""MoE token shuffling (distributed_strategy.token_shuffle_group_size): tokens are exchanged at random inside groups of
DP ranks before routing and returned afterwards.  With dropless experts and no auxiliary loss the model function is
unchanged, so loss and parameter update must equal the un-shuffled run."""
from common import *
from neuronx_distributed_training_b200.config import from_dict, process_config
from neuronx_distributed_training_b200.trainer.hf_mixtral_module import HFMixtralModule

rank, world = init()


def run(shuffle):
    c = process_config(from_dict({
        "name": "hf_mixtral", "model_source": "hf", "seed": 5,
        "trainer": {"devices": world, "num_nodes": 1, "max_steps": 10, "gradient_clip_val": 1.0},
        "distributed_strategy": {"tensor_model_parallel_size": 1, "zero1": True, "token_shuffle_group_size": shuffle},
        "data": {"micro_batch_size": 1, "global_batch_size": world},
        "model": {"model_config": "mixtral_8x7b", "encoder_seq_length": 16, "num_layers": 2, "hidden_size": 32, "intermediate_size": 48,
                  "num_attention_heads": 4, "num_kv_heads": 2, "vocab_size": 64, "fusions": {"flash_attention": False},
                  "force_cpu_initialization": True,
                  "moe": {"num_experts": 4, "top_k": 2, "dropless": True, "router_aux_loss_coef": 0.0, "glu_mlp": True},
                  "optim": {"name": "adamw", "lr": 1e-2, "weight_decay": 0.0, "betas": [0.9, 0.95]}},
        "precision": {"type": "fp32"}}), world)
    mod = HFMixtralModule(c, None)
    mod.setup()
    mod.configure_optimizers(10)
    moes = [m for m in mod.model.modules() if type(m).__name__ == "MoE"]
    assert moes and all((len(m._shuffle_ranks) == shuffle) == (shuffle > 1) for m in moes)
    g = torch.Generator().manual_seed(1)
    ids_all = torch.randint(0, 64, (world, 16), generator=g)
    dpr = ps.get_data_parallel_rank()
    losses = [float(mod.training_step({"input_ids": ids_all[dpr:dpr + 1], "labels": ids_all[dpr:dpr + 1].clone()})) for _ in range(3)]
    return losses, {n: p.detach().clone() for n, p in mod.local_named_parameters()}


l0, p0 = run(1)
l1, p1 = run(world)
for a, b in zip(l0, l1):
    assert abs(a - b) < 1e-5, (l0, l1)
for n in p0:
    check(f"param {n}", p1[n], p0[n], atol=2e-5)
finish()
