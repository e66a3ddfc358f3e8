"""Overlapped data-parallel gradient reduction (buckets reduced to their ZeRO-1 owner during the last backward):
(1) optimizer level vs torch.optim.AdamW on the full batch; (2) module level: identical losses with and without it."""
from common import *
from neuronx_distributed_training_b200.optim import DistributedOptimizer
from neuronx_distributed_training_b200.config import from_dict, process_config

rank, world = init()
dp, dpr = ps.get_data_parallel_size(), ps.get_data_parallel_rank()


def make():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(12, 33), torch.nn.Tanh(), torch.nn.Linear(33, 17), torch.nn.Tanh(), torch.nn.Linear(17, 5))


for zero1 in (True, False):
    m, ref = make(), make()
    opt = DistributedOptimizer([{"params": list(m.parameters()), "weight_decay": 0.0}], lr=1e-2, betas=(0.9, 0.95), zero1=zero1,
                               master_weights=True, fp32_grad_acc=True, max_grad_norm=0.5, dp_group=ps.get_data_parallel_group())
    assert opt.enable_grad_reduce_overlap(bucket_mb=256 * 4 / (1 << 20))          # 256-element buckets
    assert len(opt._buckets) >= 4, opt._buckets
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-2, betas=(0.9, 0.95), weight_decay=0.0)
    g = torch.Generator().manual_seed(3)
    launched_early = 0
    for step in range(4):
        opt.zero_grad()
        ropt.zero_grad()
        for mb in range(2):                                                       # two micro-batches: window only on the last
            X = torch.randn(4 * dp, 12, generator=g); Y = torch.randn(4 * dp, 5, generator=g)
            if mb == 1:
                opt.start_grad_sync_window()
            xl, yl = X.chunk(dp)[dpr], Y.chunk(dp)[dpr]
            (((m(xl) - yl) ** 2).mean() / 2).backward()
            (((ref(X) - Y) ** 2).mean() / 2).backward()
        launched_early += sum(opt._launched)                                      # buckets already in flight before step()
        opt.step()
        gn = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.5)
        ropt.step()
        check(f"grad norm zero1={zero1} step {step}", opt.grad_norm, gn, atol=1e-5)
    assert launched_early > 0, "no bucket was reduced during backward"
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        check(f"param {n} zero1={zero1}", p, q, atol=2e-5)

# ---- through the task module: same trajectory with and without the overlap ------------------------------------
from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule


def run(overlap):
    cfg = process_config(from_dict({
        "name": "t", "model_source": "hf", "seed": 7, "bucket_size_collectives": 0.05,
        "trainer": {"devices": world, "num_nodes": 1, "max_steps": 5, "gradient_clip_val": 1.0},
        "exp_manager": {"create_checkpoint_callback": False, "create_tensorboard_logger": False},
        "distributed_strategy": {"tensor_model_parallel_size": 1, "zero1": True},
        "data": {"micro_batch_size": 1, "global_batch_size": 2 * world, "synthetic": True},
        "model": {"model_config": "llama_125m", "encoder_seq_length": 32, "num_layers": 2, "hidden_size": 64, "intermediate_size": 128,
                  "num_attention_heads": 4, "num_kv_heads": 2, "vocab_size": 128, "fusions": {"flash_attention": False},
                  "use_cpu_initialization": True, "force_cpu_initialization": True,
                  "optim": {"name": "adamw_fp32OptState", "lr": 1e-2, "weight_decay": 0.01, "betas": [0.9, 0.95], "overlap_grad_reduce": overlap,
                            "sched": {"name": "LinearAnnealingWithWarmUp", "warmup_steps": 1, "max_steps": 50}}},
        "precision": {"type": "fp32"}}), world)
    mod = HFLlamaModule(cfg, None)
    mod.setup()
    mod.configure_optimizers(50)
    assert bool(getattr(mod.optimizer, "_overlap", False)) == overlap
    gen = torch.Generator().manual_seed(100 + rank)
    out = []
    for _ in range(4):
        ids = torch.randint(0, 128, (2, 32), generator=gen)
        out.append(float(mod.training_step({"input_ids": ids, "labels": ids.clone()})))
    return out, {n: p.detach().clone() for n, p in mod.local_named_parameters()}


l0, p0 = run(False)
l1, p1 = run(True)
for a, b in zip(l0, l1):
    assert abs(a - b) < 1e-6, (l0, l1)
for n in p0:
    check(f"module param {n}", p1[n], p0[n], atol=1e-6)
finish()
