"""HF-Mixtral through the 1F1B pipeline (PP=2): runs, the routers of BOTH stages receive gradients (the first stage's
auxiliary loss is attached locally), and the language-model loss tracks the single-process model (the load-balancing
term is evaluated per stage, docs/general/known_issues.md, so the comparison is approximate)."""
from common import *
from neuronx_distributed_training_b200.config import from_dict, process_config
from neuronx_distributed_training_b200.trainer.hf_mixtral_module import HFMixtralModule

rank, world = init(pp=2)


def cfg(pp):
    return process_config(from_dict({
        "name": "hf_mixtral", "model_source": "hf", "seed": 5,
        "trainer": {"devices": pp, "num_nodes": 1, "max_steps": 10, "gradient_clip_val": 1.0},
        "distributed_strategy": {"tensor_model_parallel_size": 1, "pipeline_model_parallel_size": pp, "zero1": True},
        "data": {"micro_batch_size": 1, "global_batch_size": 4},
        "model": {"model_config": "mixtral_8x7b", "encoder_seq_length": 16, "num_layers": 4, "hidden_size": 32, "intermediate_size": 48,
                  "num_attention_heads": 4, "num_kv_heads": 2, "vocab_size": 64, "fusions": {"flash_attention": False},
                  "force_cpu_initialization": True,
                  "moe": {"num_experts": 4, "top_k": 2, "dropless": True, "router_aux_loss_coef": 0.02, "glu_mlp": True},
                  "optim": {"name": "adamw", "lr": 1e-2, "weight_decay": 0.0, "betas": [0.9, 0.95]}},
        "precision": {"type": "fp32"},
    }), pp)


g = torch.Generator().manual_seed(1)
batches = [torch.randint(0, 64, (4, 16), generator=g) for _ in range(3)]
# single-process reference first (its initial weights are copied into the pipeline stages by name: global layer numbering)
ps.destroy_model_parallel()
_saved = dist.is_initialized
dist.is_initialized = lambda: False
try:
    ps.initialize_model_parallel(1, 1, 1)
    ref = HFMixtralModule(cfg(1), None); ref.setup(); ref.configure_optimizers(10)
    init_state = {n: p.detach().clone() for n, p in ref.local_named_parameters()}
    theirs = [float(ref.training_step({"input_ids": ids, "labels": ids.clone()})) for ids in batches]
finally:
    dist.is_initialized = _saved
ps.destroy_model_parallel(); ps.initialize_model_parallel(1, 2, 1)
mod = HFMixtralModule(cfg(2), None); mod.setup()
with torch.no_grad():
    for n, p in mod.local_named_parameters():
        p.copy_(init_state[n])
mod.configure_optimizers(10)
ours = []
for i, ids in enumerate(batches):
    if i == 0:
        mod.optimizer.zero_grad()
        loss = mod.forward_backward_step({"input_ids": ids, "labels": ids.clone()})
        routers = [(n, p) for n, p in mod.local_named_parameters() if "router" in n]
        assert routers, "this stage has MoE layers"
        for n, p in routers:
            gr = getattr(p, "main_grad", None)
            gr = gr if gr is not None else p.grad
            assert gr is not None and float(gr.abs().sum()) > 0, f"router {n} got no gradient on pipeline stage {ps.get_pipeline_model_parallel_rank()}"
    loss = mod.training_step({"input_ids": ids, "labels": ids.clone()})
    ours.append(float(loss))
if rank == 0:
    print("pp losses", ours, "single", theirs, flush=True)
for a, b in zip(ours, theirs):
    assert abs(a - b) < 1e-2, (ours, theirs)
finish()
