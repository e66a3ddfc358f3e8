"""Train N steps with checkpoints, then resume from the middle: identical loss trajectory and parameters."""
from common import *
import json, shutil, tempfile
from neuronx_distributed_training_b200.config import from_dict, process_config
from neuronx_distributed_training_b200.trainer import Trainer
from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
from neuronx_distributed_training_b200.data import SyntheticDataModule
from neuronx_distributed_training_b200.utils.exp_manager import exp_manager

TP = int(os.environ.get("TP", "1")); PP = int(os.environ.get("PP", "1"))
rank, world = init(tp=TP, pp=PP)
work = os.environ["WORKDIR"]
ASYNC = os.environ.get("ASYNC", "0") == "1"

def make_cfg(max_steps, log_dir):
    return from_dict({
        "name": "t", "model_source": "hf", "seed": 7,
        "trainer": {"devices": world, "num_nodes": 1, "max_steps": max_steps, "log_every_n_steps": 1, "gradient_clip_val": 1.0},
        "exp_manager": {"explicit_log_dir": log_dir, "create_tensorboard_logger": False, "resume_if_exists": True, "resume_ignore_no_checkpoint": True,
                        "create_checkpoint_callback": True, "async_checkpointing": ASYNC, "save_xser": True,
                        "checkpoint_callback_params": {"save_top_k": 2, "every_n_train_steps": 2, "save_last": False,
                                                       "filename": "t--{step}-{consumed_samples}", "monitor": "step", "mode": "max"}},
        "distributed_strategy": {"tensor_model_parallel_size": TP, "pipeline_model_parallel_size": PP, "zero1": True, "sequence_parallel": TP > 1},
        "data": {"micro_batch_size": 1, "global_batch_size": 2 * (world // (TP * PP)), "synthetic": True},
        "model": {"model_config": "llama_125m", "encoder_seq_length": 32, "num_layers": 2, "hidden_size": 64, "intermediate_size": 128,
                  "num_attention_heads": 4, "num_kv_heads": 2, "vocab_size": 128, "fusions": {"flash_attention": False},
                  "use_cpu_initialization": True, "force_cpu_initialization": True,
                  "optim": {"name": "adamw_fp32OptState", "lr": 1e-2, "weight_decay": 0.01, "betas": [0.9, 0.95],
                            "sched": {"name": "CosineAnnealing", "warmup_steps": 2, "max_steps": 8, "min_lr": 1e-4}}},
        "precision": {"type": "mixed_precision"},
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
        l = orig(b, i); losses.append(float(l)); return l
    mod.training_step = wrapped
    tr.fit(mod, dm, ckpt_path=em.resume_from_checkpoint)
    params = {n: p.detach().float().clone() for n, p in mod.local_named_parameters()}
    import logging
    for h in list(logging.getLogger("nxdt").handlers): logging.getLogger("nxdt").removeHandler(h)
    return losses, params, tr

full_dir, res_dir = os.path.join(work, "full"), os.path.join(work, "resumed")
l_full, p_full, _ = run(8, full_dir)
l_a, _, _ = run(4, res_dir)                       # stops at step 4 (checkpoint at steps 2 and 4)
ck = sorted(os.listdir(os.path.join(res_dir, "checkpoints")))
assert any("step=4" in c for c in ck), ck
root = os.path.join(res_dir, "checkpoints", [c for c in ck if "step=4" in c][0])
for f in ("done", "checkpoint", "user_content.pt", "model", "optim"):
    assert os.path.exists(os.path.join(root, f)), f
if rank == 0:
    assert os.path.exists(os.path.join(root, "model", f"dp_rank_00_tp_rank_00_pp_rank_00.pt.tensors")), os.listdir(os.path.join(root, "model"))
l_b, p_res, tr = run(8, res_dir)                  # auto-resume from step 4, continue to 8
assert tr.global_step == 8 and len(l_b) == 4, (tr.global_step, len(l_b))
for a, b in zip(l_full[4:], l_b):
    assert abs(a - b) < 1e-6, (l_full, l_a, l_b)
for n in p_full:
    check(f"param {n}", p_res[n], p_full[n], atol=1e-6)
ck = sorted(os.listdir(os.path.join(res_dir, "checkpoints")))
assert len([c for c in ck if c.endswith(".ckpt")]) == 2, ck       # save_top_k=2 pruned the rest
finish()
