#!/usr/bin/env python
"""Headline benchmark: Llama-3-8B training tokens/s (whole job, device-timed, max over ranks).

    python bench.py --gpus N --steps K --warmup W            # N=1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Config = BASELINE.json #2: ``hf_llama3_8B_config`` (Llama-3-8B, seq 8192, bf16 compute, fp32 master
weights / fp32 grad accumulation, AdamW fp32 state, TP=N with sequence parallelism, micro batch 1),
synthetic tokens, random-init weights.  The global batch is fixed (strong scaling).
One optimizer step = ``global_batch`` sequences; every timed step includes forward, backward, grad-norm
clipping, the AdamW update and the LR schedule.

Other named configs: ``--model llama2_7b --tp 2 --pp 2`` on 8 GPUs (BASELINE config 3: TP2 x PP2 x DP2, 1F1B, ZeRO-1),
``--model llama3_70b`` on 8 GPUs (config 5).  Experiment switches (all default off): ``NXDT_CUDA_GRAPH=1`` (graph replay of
the micro-batch), ``NXDT_MB_OVERLAP=1`` (two micro-batches on two streams, NCCL TP path), ``NXDT_FUSED_TP=0`` (NCCL TP
path), ``NXDT_OVERLAP_GRAD_REDUCE=1`` (data-parallel reduction during backward), ``NXDT_FP8=1`` (e4m3 linear GEMMs),
``NXDT_KEEP_GATHERED=1`` (column linears keep the gathered input for wgrad: one all-gather less per linear).

``--impl reference`` would run the unmodified reference from baseline/_ref; it cannot be imported here
(needs torch_xla / neuronx_distributed / nemo / lightning) so that arm reports unavailable.
``--impl baseline`` runs this framework with every hot op on the library path (NCCL + cuBLAS + SDPA).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the 1-GPU run keeps 144 GB of persistent state on a 180 GB part: growable segments avoid fragmentation OOMs
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "baseline", "reference"])
    ap.add_argument("--global-batch", type=int, default=int(os.environ.get("NXDT_BENCH_GBS", "8")))
    ap.add_argument("--seq-len", type=int, default=8192)
    ap.add_argument("--layers", type=int, default=0, help="debug only (0 = the model's own depth, which is what the "
                                                          "reported config uses)")
    ap.add_argument("--model", default="llama3_8b")
    ap.add_argument("--recompute", default=os.environ.get("NXDT_BENCH_RECOMPUTE", "auto"))
    ap.add_argument("--tp", type=int, default=0, help="tensor-parallel size (default: all GPUs, the headline config)")
    ap.add_argument("--pp", type=int, default=1, help="pipeline-parallel size (BASELINE config 3: --model llama2_7b "
                                                      "--tp 2 --pp 2 on 8 GPUs)")
    ap.add_argument("--cp", type=int, default=1, help="context-parallel size (long-context runs: --seq-len 32768 --tp 4 --cp 2)")
    ap.add_argument("--fp8", action="store_true", help="BASELINE config 4: linear-layer GEMMs on the e4m3 tcgen05 kernel "
                    "(model.fusions.fp8_gemm; 1xK row / column block scales, fp32 accumulation)")
    ap.add_argument("--nvls", action="store_true", help="model.fusions.nvls_norm_allgather: RMSNorm + sequence-parallel "
                    "all-gather as one NVSwitch-multicast kernel for the linears the fused AG->GEMM kernel leaves on NCCL")
    ap.add_argument("--overlap", action="store_true", help="two-stream micro-batches on the NCCL tensor-parallel path "
                    "(model.fusions.microbatch_overlap): collectives of one micro-batch overlap the GEMMs of the other")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def reference_unavailable():
    why = ("reference installs offline only with --no-deps and cannot be imported: it requires torch_xla, "
           "neuronx_distributed, nemo_toolkit, lightning, hydra (none in the image, no network); it has no CUDA path")
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    try:
        sys.path.insert(0, ref_dir)
        import neuronx_distributed_training  # noqa: F401
        why = "reference imported unexpectedly but has no CUDA/NCCL execution path (XLA backend only)"
    except Exception as e:  # expected
        why += f" [{type(e).__name__}: {str(e)[:80]}]"
    finally:
        if ref_dir in sys.path:
            sys.path.remove(ref_dir)
    print(json.dumps({"impl": "reference", "unavailable": why}))


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # median over the samples taken under load (power above idle) when we can tell them apart
        load = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def _keep_gathered():
    from neuronx_distributed_training_b200.parallel.layers import keep_gathered_input
    return keep_gathered_input()


def main():
    args = parse()
    if args.impl == "reference":
        reference_unavailable()
        return
    import torch
    import torch.distributed as dist
    from neuronx_distributed_training_b200 import ops
    from neuronx_distributed_training_b200.config import from_dict, process_config
    from neuronx_distributed_training_b200.parallel import state as ps
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    from neuronx_distributed_training_b200.data.base import SyntheticTokenDataset, default_collate, pin_batch

    if args.overlap:
        os.environ["NXDT_MB_OVERLAP"] = "1"
        os.environ.setdefault("NCCL_MAX_CTAS", "16")      # the collectives live in the SMs the capped GEMM grid leaves free
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun --nproc-per-node {args.gpus}"
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
    if world > 1:
        ps.init_distributed("nccl")
    tp = args.tp or world
    pp = args.pp
    cp = args.cp
    assert world % (tp * pp * cp) == 0, f"world {world} not divisible by tp*pp*cp = {tp * pp * cp}"
    dp = world // (tp * pp * cp)
    ps.initialize_model_parallel(tensor_model_parallel_size=tp, pipeline_model_parallel_size=pp, context_parallel_size=cp)
    if args.impl == "baseline":
        ops.set_backend("torch")
    recompute = args.recompute
    if recompute == "auto":
        # single GPU holds the full fp32 optimizer state (144 GB): selective recomputation of the SwiGLU
        # activation (the reference's `ActivationMultiplyMLP` recompute class); with TP >= 2 the per-rank
        # footprint needs no recompute.
        # (the library-path arm keeps whole-MLP recomputation: its unfused ops hold more activations)
        recompute = ("swiglu" if args.impl == "native" else "mlp") if world == 1 else "none"
        if world > 1:
            # larger models (config 5: Llama-3-70B at TP 8 keeps 8.8 B parameters + fp32 AdamW state per GPU): take the
            # lightest recompute policy whose estimated footprint fits (utils/perf_estimate.py, calibrated on the 8B runs)
            import json as _json
            from neuronx_distributed_training_b200.utils.perf_estimate import estimate_memory_gib
            cpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "neuronx_distributed_training_b200", "models",
                                 "configs", args.model + ".json")
            if os.path.exists(cpath):
                c = _json.load(open(cpath))
                for cand in ("none", "swiglu", "mlp", "full"):
                    est = estimate_memory_gib(
                        c["hidden_size"], args.layers or c["num_hidden_layers"], args.seq_len, c["vocab_size"], c["intermediate_size"],
                        c["num_attention_heads"], c.get("num_key_value_heads"), c.get("head_dim"), tp=tp, pp=pp, dp=dp,
                        recompute=None if cand == "none" else cand)
                    recompute = cand
                    if est["total"] <= (165.0 if args.impl == "native" else 150.0):
                        break
    gran, rec_list = {"none": (None, None), "mlp": ("selective", ["CoreAttention", "LlamaMLP"]),
                      "mlp_only": ("selective", ["LlamaMLP"]),
                      "swiglu": ("selective", ["ActivationMultiplyMLP"]),
                      "swiglu_norm": ("selective", ["ActivationMultiplyMLP", "LlamaRMSNorm"]),
                      "full": ("full", None)}[recompute]
    cfg = from_dict({
        "name": "hf_llama", "model_source": "hf", "seed": 1234,
        "trainer": {"devices": world, "num_nodes": 1, "max_steps": 100000, "log_every_n_steps": 1000000,
                    "gradient_clip_val": 1.0},
        "exp_manager": {"create_checkpoint_callback": False, "create_tensorboard_logger": False},
        "distributed_strategy": {"tensor_model_parallel_size": tp, "pipeline_model_parallel_size": pp,
                                 "context_parallel_size": cp, "zero1": True, "sequence_parallel": tp > 1, "kv_replicator": 1},
        "data": {"micro_batch_size": 1, "global_batch_size": args.global_batch, "synthetic": True},
        "model": {"model_config": args.model, "encoder_seq_length": args.seq_len,
                  "max_position_embeddings": args.seq_len, **({"num_layers": args.layers} if args.layers else {}),
                  "qkv_linear": True,
                  "fuse_qkv": True, "activations_checkpoint_granularity": gran,
                  "activations_checkpoint_recompute": rec_list,
                  "fusions": {"flash_attention": cp == 1, "ring_attention": cp > 1, "linear_cross_entropy": True,
                              "fp8_gemm": bool(args.fp8), "nvls_norm_allgather": bool(args.nvls)},
                  "optim": {"name": "adamw_fp32OptState", "lr": 1.5e-4, "weight_decay": 0.01, "betas": [0.9, 0.999],
                            "sched": {"name": "LinearAnnealingWithWarmUp", "warmup_steps": 100, "max_steps": 100000}}},
        "precision": {"type": "mixed_precision"},
    })
    process_config(cfg, world)
    module = HFLlamaModule(cfg, None)
    module.setup("fit")
    module.configure_optimizers(100000)
    # the fused TP kernels are attached by module.setup() itself (model.fusions.fused_tp: auto; NXDT_FUSED_TP=0 or the
    # library baseline arm select the NCCL path)
    if rank == 0 and tp > 1:
        print(f"[bench] fused TP kernels on {getattr(module, 'fused_tp_layers', 0)} linears", file=sys.stderr)
    hf = module.hf_config
    tokens_per_step = args.global_batch * args.seq_len

    ds = SyntheticTokenDataset(hf.vocab_size if hf.vocab_size <= 128256 else 128256, args.seq_len, 1 << 20, 1234)

    def host_batch(step):
        b = default_collate([ds[step * args.global_batch + i] for i in range(args.global_batch)])
        return pin_batch(b)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (untimed) ----
    nb = args.warmup + 2 * args.steps + 1
    batches = [host_batch(i) for i in range(nb)]
    it = 0
    for _ in range(args.warmup):
        loss = module.training_step(batches[it]); it += 1
    barrier()
    mem_gb = torch.cuda.max_memory_allocated() / 2 ** 30

    # ---- timed region 1: device-timed, K steps, inputs pre-staged on the device, no host readback ----
    dev_batches = [{k: v.cuda(non_blocking=True) for k, v in batches[it + i].items()} for i in range(args.steps)]
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ops.reset_native_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        loss = module.training_step(dev_batches[i])
    ev1.record()
    barrier()
    launches = ops.native_launch_count()
    dev_ms = ev0.elapsed_time(ev1)
    it += args.steps
    del dev_batches

    # ---- timed region 2: end to end through the public step API: pinned-host inputs -> H2D, loss -> host ----
    e2e = None
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        h2d = d2h = 0
        last = None
        for i in range(args.steps):
            b = batches[it + i]
            h2d += sum(v.numel() * v.element_size() for v in b.values() if torch.is_tensor(v))
            loss = module.training_step(b)
            last = float(loss)            # device -> host read of the step result
            d2h += loss.numel() * loss.element_size()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()

    # ---- reduce: max over ranks ----
    t = torch.tensor([dev_ms, e2e_ms if not args.no_e2e else 0.0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        value = tokens_per_step * args.steps / (dev_ms / 1e3)
        out = {
            "metric": ("Llama-3-8B" if args.model == "llama3_8b" else args.model)
                      + " training tokens/sec (whole job, device-timed, max over ranks)",
            "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": "fp8 e4m3 linear GEMMs (1xK block scales, fp32 accumulate), bf16 elsewhere" if args.fp8 else "bf16", "data": "synthetic tokens, random-init weights",
            "impl": args.impl,
            "config": {"model": (args.model.replace("_", "-") if args.layers in (0, hf.num_hidden_layers)
                                 else f"{args.model.replace('_', '-')}-{args.layers}L"),
                       "global_batch": args.global_batch, "micro_batch": 1,
                       "seq_len": args.seq_len,
                       "parallelism": f"tp{tp}" + ("+sp" if tp > 1 else "") + (f"+pp{pp}" if pp > 1 else "")
                                      + (f"+cp{cp}" if cp > 1 else "") + (f"+dp{dp}(zero1)" if dp > 1 else ""),
                       "precision": "bf16 compute, fp32 master weights + fp32 grad accumulation, AdamW fp32 state",
                       "recompute": recompute, "l2_policy": "inputs larger than L2 (weights+activations >> 126 MB per step)",
                       "peak_mem_gb": round(mem_gb, 1), "final_loss": float(loss),
                       "cuda_graph": getattr(module, "_mb_graph", None) is not None,
                       "microbatch_overlap": bool(getattr(module, "use_mb_overlap", False)),
                       "fused_tp_layers": int(getattr(module, "fused_tp_layers", 0)),
                       "fused_zero1": bool(getattr(module, "fused_zero1", False)),
                       "keep_gathered_input": bool(_keep_gathered()),
                       "nvls_norm_allgather_layers": int(getattr(module, "nvls_layers", 0))},
            "clocks": clocks, "gpu_launches": launches,
        }
        if not args.no_e2e:
            out["e2e"] = {"value": tokens_per_step * args.steps / (e2e_ms / 1e3), "unit": "tokens/s",
                          "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                          "ms_per_step": e2e_ms / args.steps}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        if getattr(module, "_mb_graph", None) is not None:
            # a captured graph holds NCCL kernels of this communicator: drop it before the group is destroyed
            module._mb_graph = None
            import gc
            gc.collect()
            torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
