"""Whole-model parity ON HARDWARE: Llama in bf16 under TP=N (+SP) with the fused AG->GEMM / GEMM->RS kernels, the native
flash attention and the fused LM-head loss, against the same weights on one GPU (TP=1).  Every parameter gradient is
compared per tensor: relative loss error <= 1e-3, cosine >= 0.999, norm within 2 %.  (The CPU twin is
tests/dist_scripts/tp_model.py, fp32 / gloo / allclose.)    torchrun --nproc-per-node N check_tp_model_gpu.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from neuronx_distributed_training_b200 import ops
from neuronx_distributed_training_b200.parallel import state as ps
from neuronx_distributed_training_b200.parallel.fused import enable_fused_tp
from neuronx_distributed_training_b200.models.llama import LlamaConfig, LlamaForCausalLM

ps.init_distributed("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
TP = world
dev = torch.device("cuda", torch.cuda.current_device())
S = int(os.environ.get("SEQ", "2048"))
FUSED = os.environ.get("FUSED", "1") == "1"
NVLS = os.environ.get("NVLS", "0") == "1"     # RMSNorm + all-gather as one NVSwitch-multicast kernel (parallel/nvls.py)
NVLS_KEEP = os.environ.get("NVLS_KEEP", "1") == "1"
base = dict(vocab_size=8192, hidden_size=1024, intermediate_size=2816, num_hidden_layers=int(os.environ.get("LAYERS", "3")),
            num_attention_heads=8, num_key_value_heads=max(2, TP), head_dim=128, torch_dtype="bfloat16",
            use_cpu_initialization=True)
g = torch.Generator().manual_seed(5)
ids = torch.randint(0, 8192, (2, S), generator=g).to(dev)


def build(tp_size, sp):
    ps.destroy_model_parallel()
    ps.initialize_model_parallel(tp_size, 1, 1)
    torch.manual_seed(11)
    return LlamaForCausalLM(LlamaConfig(**base, sequence_parallel_enabled=sp, qkv_linear=True, fuse_qkv=True), device=dev)


def fwd_bwd(m, tp_size):
    n_fused = enable_fused_tp(m, seq_len=S, micro_batch=2) if (FUSED and tp_size > 1) else 0
    if NVLS and tp_size > 1:
        from neuronx_distributed_training_b200.parallel.nvls import enable_nvls_norm_allgather
        from neuronx_distributed_training_b200.parallel.layers import set_keep_gathered_input
        n_nvls = enable_nvls_norm_allgather(m, seq_len=S, micro_batch=2)
        if n_nvls == 0:
            if rank == 0:
                print("NVLS_UNAVAILABLE", flush=True)
            dist.barrier(); dist.destroy_process_group(); sys.exit(0)
        set_keep_gathered_input(NVLS_KEEP)
        print(f"[rank {rank}] NVLS norm + all-gather attached to {n_nvls} linears (keep gathered input: {NVLS_KEEP})", flush=True)
    ops.reset_native_launch_count()
    loss, _ = m(input_ids=ids, labels=ids)
    loss.backward()
    torch.cuda.synchronize()
    grads = {}
    for n, p in m.named_parameters():
        gr = (p.main_grad if getattr(p, "main_grad", None) is not None else p.grad).float().clone()
        if tp_size > 1 and getattr(p, "sequence_parallel_enabled", False):
            dist.all_reduce(gr)                  # each TP rank saw S/TP of the tokens (the optimizer does this sum)
        grads[n] = gr
    return loss.detach().float().clone(), grads, n_fused, ops.native_launch_count()


# The SAME weights in both layouts: the TP=1 model's state goes through the checkpoint converter (framework -> full HF
# naming -> TP shards), i.e. the code path a user takes to re-shard a checkpoint; gradients are compared the same way.
from neuronx_distributed_training_b200.checkpoint.converter import CheckpointConverterBase
ccfg = {"num_attention_heads": base["num_attention_heads"], "num_key_value_heads": base["num_key_value_heads"],
        "hidden_size": base["hidden_size"], "head_dim": base["head_dim"], "num_hidden_layers": base["num_hidden_layers"]}
conv1 = CheckpointConverterBase(tp_size=1, pp_size=1, model_style="hf", config=ccfg, kv_size_multiplier=1, fuse_qkv=True, qkv_linear=True)
convN = CheckpointConverterBase(tp_size=TP, pp_size=1, model_style="hf", config=ccfg, kv_size_multiplier=1, fuse_qkv=True, qkv_linear=True)
ref = build(1, False)
loss_1, g_1, _, _ = fwd_bwd(ref, 1)
full_w = conv1.unshard([[{n: p.detach() for n, p in ref.named_parameters()}]])
full_g = conv1.unshard([[g_1]])
del ref
m = build(TP, True)
tpr = ps.get_tensor_model_parallel_rank()
missing = m.load_state_dict(convN.shard_for_rank(full_w, tpr, 0), strict=False)
assert not missing.missing_keys, missing.missing_keys
loss_tp, g_tp, n_fused, launches = fwd_bwd(m, TP)
want_g = convN.shard_for_rank(full_g, tpr, 0)
bad = []
rel_loss = abs(loss_tp.item() - loss_1.item()) / abs(loss_1.item())
if rel_loss > 1e-3:
    bad.append(f"loss {loss_tp.item()} vs {loss_1.item()}")
worst_cos, worst_norm = 1.0, 0.0
assert set(g_tp) == set(want_g), sorted(set(g_tp) ^ set(want_g))[:6]
for n, gr in g_tp.items():
    want = want_g[n].to(gr.device).float()
    cos = torch.nn.functional.cosine_similarity(gr.flatten(), want.flatten(), dim=0).item()
    nr = (gr.norm() / want.norm().clamp_min(1e-12)).item()
    worst_cos, worst_norm = min(worst_cos, cos), max(worst_norm, abs(nr - 1))
    if cos < 0.999 or abs(nr - 1) > 0.02:
        bad.append(f"{n}: cosine {cos:.5f} norm ratio {nr:.4f}")
flag = torch.tensor([len(bad)], device=dev)
dist.all_reduce(flag)
print(f"[rank {rank}] TP={TP} fused layers {n_fused} native launches {launches}: loss {loss_tp.item():.5f} vs TP=1 {loss_1.item():.5f} "
      f"(rel {rel_loss:.2e}); worst cosine {worst_cos:.5f}, worst |norm ratio - 1| {worst_norm:.4f}" + ("".join("\n   BAD " + b for b in bad[:8])),
      flush=True)
if FUSED:
    assert n_fused > 0, "the fused tensor-parallel kernels were not attached"
assert launches > 0, "no native kernels were launched"
dist.barrier()
if rank == 0:
    print("TP_MODEL_GPU_OK" if flag.item() == 0 else "TP_MODEL_GPU_MISMATCH", flush=True)
dist.destroy_process_group()
