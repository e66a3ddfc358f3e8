"""Phase of the DP-resharding test (one phase per launch, the world size differs between phases):
PHASE=save   : train STEPS_A steps with the launch's DP size, checkpointing (ZeRO-1 optimizer shards per DP rank)
PHASE=resume : auto-resume from that directory with THIS launch's DP size and continue to STEPS_B
PHASE=full   : uninterrupted STEPS_B steps
Every phase writes its loss list to $WORKDIR/<phase>_losses.json (rank 0)."""
from common import *
import json
from neuronx_distributed_training_b200.config import from_dict, process_config
from neuronx_distributed_training_b200.trainer import Trainer
from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
from neuronx_distributed_training_b200.data import SyntheticDataModule
from neuronx_distributed_training_b200.utils.exp_manager import exp_manager

rank, world = init(tp=1, pp=1)
work, phase = os.environ["WORKDIR"], os.environ["PHASE"]
steps_a, steps_b = int(os.environ.get("STEPS_A", "4")), int(os.environ.get("STEPS_B", "8"))


def make_cfg(max_steps, log_dir):
    return from_dict({
        "name": "t", "model_source": "hf", "seed": 7,
        "trainer": {"devices": world, "num_nodes": 1, "max_steps": max_steps, "log_every_n_steps": 1, "gradient_clip_val": 1.0},
        "exp_manager": {"explicit_log_dir": log_dir, "create_tensorboard_logger": False, "resume_if_exists": True,
                        "resume_ignore_no_checkpoint": True, "create_checkpoint_callback": True, "save_xser": True,
                        "checkpoint_callback_params": {"save_top_k": 1, "every_n_train_steps": steps_a, "save_last": False,
                                                       "filename": "t--{step}-{consumed_samples}", "monitor": "step", "mode": "max"}},
        "distributed_strategy": {"tensor_model_parallel_size": 1, "pipeline_model_parallel_size": 1, "zero1": True},
        "data": {"micro_batch_size": 1, "global_batch_size": 4, "synthetic": True},
        "model": {"model_config": "llama_125m", "encoder_seq_length": 32, "num_layers": 2, "hidden_size": 64, "intermediate_size": 128,
                  "num_attention_heads": 4, "num_kv_heads": 2, "vocab_size": 128, "fusions": {"flash_attention": False},
                  "use_cpu_initialization": True, "force_cpu_initialization": True,
                  "optim": {"name": "adamw_fp32OptState", "lr": 1e-2, "weight_decay": 0.01, "betas": [0.9, 0.95],
                            "sched": {"name": "CosineAnnealing", "warmup_steps": 2, "max_steps": 8, "min_lr": 1e-4}}},
        "precision": {"type": "fp32"},
    })


def run(max_steps, log_dir):
    cfg = process_config(make_cfg(max_steps, log_dir), world)
    tr = Trainer(**cfg.trainer.to_dict())
    em = exp_manager(tr, cfg.exp_manager)
    dm = SyntheticDataModule(cfg, tr)
    mod = HFLlamaModule(cfg, tr)
    losses = []
    orig = mod.training_step

    def wrapped(b, i=0):
        l = orig(b, i)
        losses.append(float(l))
        return l
    mod.training_step = wrapped
    tr.fit(mod, dm, ckpt_path=em.resume_from_checkpoint)
    return losses, tr


if phase == "save":
    losses, tr = run(steps_a, os.path.join(work, "run"))
    assert tr.global_step == steps_a
elif phase == "resume":
    losses, tr = run(steps_b, os.path.join(work, "run"))
    assert tr.global_step == steps_b and len(losses) == steps_b - steps_a, (tr.global_step, len(losses))
else:
    losses, tr = run(steps_b, os.path.join(work, "full"))
if rank == 0:
    with open(os.path.join(work, f"{phase}_losses.json"), "w") as f:
        json.dump(losses, f)
finish()
