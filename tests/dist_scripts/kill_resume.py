"""Fault injection: a rank dies hard (os._exit) in the middle of training; the relaunched job auto-resumes from the
newest complete checkpoint and ends with exactly the parameters of an uninterrupted run."""
from common import *
import json
from neuronx_distributed_training_b200.config import from_dict, process_config
from neuronx_distributed_training_b200.trainer import Trainer
from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
from neuronx_distributed_training_b200.data import SyntheticDataModule
from neuronx_distributed_training_b200.utils.exp_manager import exp_manager

rank, world = init()
work, phase = os.environ["WORKDIR"], os.environ["PHASE"]       # phase: full | crash | resume
kill_at = int(os.environ.get("KILL_AT", "0"))

def cfg(log_dir):
    return from_dict({
        "name": "t", "model_source": "hf", "seed": 11,
        "trainer": {"devices": world, "num_nodes": 1, "max_steps": 6, "log_every_n_steps": 1, "gradient_clip_val": 1.0},
        "exp_manager": {"explicit_log_dir": log_dir, "create_tensorboard_logger": False, "resume_if_exists": True, "resume_ignore_no_checkpoint": True,
                        "checkpoint_callback_params": {"save_top_k": 3, "every_n_train_steps": 2, "save_last": False, "filename": "t--{step}-{consumed_samples}"}},
        "distributed_strategy": {"tensor_model_parallel_size": 1, "zero1": True},
        "data": {"micro_batch_size": 1, "global_batch_size": 2 * world, "synthetic": True},
        "model": {"model_config": "llama_125m", "encoder_seq_length": 32, "num_layers": 2, "hidden_size": 64, "intermediate_size": 128, "num_attention_heads": 4,
                  "num_kv_heads": 2, "vocab_size": 128, "fusions": {"flash_attention": False}, "force_cpu_initialization": True,
                  "optim": {"name": "adamw_fp32OptState", "lr": 1e-2, "weight_decay": 0.01, "betas": [0.9, 0.95]}},
        "precision": {"type": "mixed_precision"}})

log_dir = os.path.join(work, "full" if phase == "full" else "faulty")
c = process_config(cfg(log_dir), world)
tr = Trainer(**c.trainer.to_dict()); em = exp_manager(tr, c.exp_manager)
mod = HFLlamaModule(c, tr)
if phase == "crash":
    orig = mod.training_step
    def step(b, i=0):
        if tr.global_step == kill_at and rank == world - 1:
            os._exit(137)                      # simulated node failure: no cleanup, no checkpoint
        return orig(b, i)
    mod.training_step = step
tr.fit(mod, SyntheticDataModule(c, tr), ckpt_path=em.resume_from_checkpoint)
params = torch.cat([p.detach().float().reshape(-1) for _, p in mod.local_named_parameters()])
torch.save(params, os.path.join(work, f"{phase}_rank{rank}.pt"))
if phase == "resume":
    assert tr.global_step == 6 and tr.steps_since_resume == 2, (tr.global_step, tr.steps_since_resume)   # resumed from step 4
    ref = torch.load(os.path.join(work, f"full_rank{rank}.pt"))
    check("params after kill+resume", params, ref, atol=1e-6)
finish()
