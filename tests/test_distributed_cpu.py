"""Multi-process gloo tests (world_size > 1 on the CPU box)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
# a separate port range per pytest-xdist worker (gw0, gw1, ...) so that `-n 4` runs do not collide
_PORT = [29600 + 200 * int((os.environ.get("PYTEST_XDIST_WORKER", "gw0")[2:] or 0))]


def run_dist(script, nproc, env=None, timeout=600):
    _PORT[0] += 1
    e = dict(os.environ)
    e.update({"OMP_NUM_THREADS": "1"})
    e.update(env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_PORT[0]), os.path.join(HERE, "dist_scripts", script)]
    r = subprocess.run(cmd, env=e, capture_output=True, text=True, timeout=timeout, cwd=os.path.join(HERE, "dist_scripts"))
    assert r.returncode == 0 and "DIST_TEST_OK" in r.stdout, f"{script} failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"


pytestmark = pytest.mark.dist


@pytest.mark.parametrize("tp", [2, 4])
def test_tp_sp_layers_match_dense(tp):
    run_dist("tp_layers.py", tp, {"TP": str(tp)})


@pytest.mark.parametrize("tp", [2, 4])
def test_tensor_parallel_model_matches_single_process_gradients(tp):
    """Every parameter gradient of a TP(+SP) Llama equals the single-process one (caught: d(hidden) of the fused LM-head
    loss used to be summed over the TP group twice)."""
    run_dist("tp_model.py", tp, {"TP": str(tp)})


def test_norm_inside_column_linear_path_matches_norm_then_linear():
    """parallel/nvls.py call path (RMSNorm + all-gather fused into the column linear's autograd node), emulated workspace."""
    run_dist("nvls_pre_norm.py", 2, {"TP": "2"})


def test_megatron_gpt_tensor_parallel_matches_single_process_gradients():
    run_dist("megatron_tp_model.py", 2, {"TP": "2"})


def test_mixtral_tensor_parallel_matches_single_process_gradients():
    run_dist("mixtral_tp_model.py", 2, {"TP": "2"})


def test_lora_adapters_under_tensor_parallelism_match_single_process():
    run_dist("lora_tp_model.py", 2, {"TP": "2"})


@pytest.mark.parametrize("tp,nproc", [(2, 2), (4, 4), (2, 4)])
def test_task_module_step_under_tensor_parallelism_matches_single_process(tp, nproc):
    """tp=4 with 2 KV heads also exercises KV-head replication (kv_replicator=2) and the replica-gradient sync;
    (tp=2, 4 ranks) adds data parallelism with the ZeRO-1 sharded optimizer."""
    run_dist("tp_trainer_step.py", nproc, {"TP": str(tp)})


@pytest.mark.parametrize("tp,kvm", [(2, 1), (4, 2)])
def test_transformers_weights_through_the_converter_into_tensor_parallel_model(tp, kvm):
    """(4, 2): TP larger than the number of KV heads -- the converter replicates every KV head (kv_size_multiplier)."""
    pytest.importorskip("transformers")
    run_dist("hf_weights_tp.py", tp, {"TP": str(tp), "KVM": str(kvm)})


def test_transformers_mixtral_weights_through_the_converter_into_tensor_parallel_model():
    pytest.importorskip("transformers")
    run_dist("hf_mixtral_weights_tp.py", 2, {"TP": "2"})


def test_zero1_matches_adamw():
    run_dist("zero1.py", 2)


def test_zero1_adafactor_matches_unsharded_reference():
    pytest.importorskip("transformers")
    run_dist("zero1_adafactor.py", 2)


@pytest.mark.parametrize("pp,vpp,tp,nmb", [(2, 1, 1, 4), (4, 1, 1, 6), (2, 2, 1, 4), (2, 1, 2, 3), (4, 2, 1, 8)])
def test_pipeline_matches_unpipelined(pp, vpp, tp, nmb):
    run_dist("pipeline.py", pp * tp, {"PP": str(pp), "VPP": str(vpp), "TP": str(tp), "NMB": str(nmb)})


@pytest.mark.parametrize("tp,pp,nproc,async_", [(1, 1, 2, "0"), (2, 1, 2, "0"), (1, 2, 2, "0"), (1, 1, 2, "1")])
def test_checkpoint_resume_bit_exact(tmp_path, tp, pp, nproc, async_):
    run_dist("ckpt_resume.py", nproc, {"TP": str(tp), "PP": str(pp), "WORKDIR": str(tmp_path), "ASYNC": async_})


@pytest.mark.parametrize("cp", [2, 4])
def test_ring_attention_context_parallel(cp):
    run_dist("ring_attention.py", cp, {"CP": str(cp)})


@pytest.mark.parametrize("cp", [2, 4])
def test_ulysses_attention_context_parallel(cp):
    run_dist("ulysses_attention.py", cp, {"CP": str(cp)})


def test_checkpoint_load_reads_once_per_model_shard_and_broadcasts(tmp_path):
    run_dist("ckpt_broadcast_load.py", 2, {"WORKDIR": str(tmp_path)})


def test_context_parallel_composes_with_tensor_parallel():
    run_dist("cp_tp_model.py", 4)


def test_megatron_gpt_through_the_pipeline():
    run_dist("megatron_pp.py", 2)


def test_mixtral_through_the_pipeline():
    run_dist("mixtral_pp.py", 2)


def test_moe_expert_parallel_matches_single_process():
    run_dist("moe_ep.py", 2, {"EP": "2"})


def test_moe_expert_parallel_with_tensor_parallel_matches_single_process():
    """EP=2 x TP=2 (+SP) on 4 ranks: all-to-all dispatch of tokens to FFN-sharded experts, one optimizer step."""
    run_dist("moe_ep.py", 4, {"EP": "2", "TP": "2"})


def test_kill_and_resume_fault_injection(tmp_path):
    """SURVEY section 4 fault-injection tier: hard-kill a rank mid-run, relaunch, identical end state."""
    env = {"WORKDIR": str(tmp_path)}
    run_dist("kill_resume.py", 2, {**env, "PHASE": "full"})
    _PORT[0] += 1
    e = dict(os.environ); e.update({"OMP_NUM_THREADS": "1", **env, "PHASE": "crash", "KILL_AT": "5", "TORCH_NCCL_ASYNC_ERROR_HANDLING": "1"})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_PORT[0]), os.path.join(HERE, "dist_scripts", "kill_resume.py")]
    r = subprocess.run(cmd, env=e, capture_output=True, text=True, timeout=600, cwd=os.path.join(HERE, "dist_scripts"))
    assert r.returncode != 0                                   # the job really died
    ck = os.listdir(tmp_path / "faulty" / "checkpoints")
    assert any("step=4" in c for c in ck) and not any("step=6" in c for c in ck)
    run_dist("kill_resume.py", 2, {**env, "PHASE": "resume"})


_RESHARD_CACHE = {}


@pytest.mark.parametrize("dp_save,dp_resume", [(2, 1), (1, 2), (2, 4)])
def test_resume_with_a_different_data_parallel_size(tmp_path, dp_save, dp_resume):
    """ZeRO-1 optimizer shards written by dp_save ranks are re-assembled for dp_resume ranks; the continued loss
    trajectory matches an uninterrupted run (same global batch, only the reduction order differs)."""
    import json
    env = {"WORKDIR": str(tmp_path)}
    run_dist("reshard_dp.py", dp_save, {**env, "PHASE": "save"})
    run_dist("reshard_dp.py", dp_resume, {**env, "PHASE": "resume"})
    if "full" not in _RESHARD_CACHE:            # the uninterrupted single-process reference is the same for every pair
        run_dist("reshard_dp.py", 1, {**env, "PHASE": "full"})
        _RESHARD_CACHE["full"] = json.load(open(tmp_path / "full_losses.json"))
    full = _RESHARD_CACHE["full"]
    saved = json.load(open(tmp_path / "save_losses.json"))
    resumed = json.load(open(tmp_path / "resume_losses.json"))
    assert len(saved) == 4 and len(resumed) == 4 and len(full) == 8
    for a, b in zip(full, saved + resumed):
        assert abs(a - b) < 2e-5, (full, saved, resumed)


@pytest.mark.parametrize("dp", [2, 4])
def test_overlapped_gradient_reduction_matches(dp):
    run_dist("zero1_overlap.py", dp)


def test_moe_token_shuffling_is_transparent():
    run_dist("moe_shuffle.py", 2)
