"""The shell entry point end to end on the CPU box: train.sh -> torchrun (2 gloo ranks) -> orchestrator -> trainer."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.dist

TINY = ["model.num_layers=2", "model.hidden_size=64", "model.encoder_seq_length=32", "+model.intermediate_size=128",
        "+model.num_attention_heads=4", "+model.num_kv_heads=2", "+model.vocab_size=128"]


def _run(tmp_path, env_extra, overrides, port):
    env = dict(os.environ)
    env.update({"CONF_FILE": "hf_llama_125M_cpu_config", "LOG_PATH": str(tmp_path), "EXPLICIT_LOGDIR": str(tmp_path / "exp"),
                "MASTER_PORT": str(port), "OMP_NUM_THREADS": "1"})
    env.update(env_extra)
    r = subprocess.run(["bash", "train.sh"] + TINY + overrides, cwd=os.path.join(ROOT, "examples"), env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_train_sh_runs_the_cpu_config_and_resumes(tmp_path):
    out = _run(tmp_path, {"TRAIN_ITERS": "4"}, ["exp_manager.checkpoint_callback_params.every_n_train_steps=2"], 29781)
    steps = [int(m) for m in re.findall(r"step (\d+) \| reduced_train_loss", out)]
    assert steps == [1, 2, 3, 4], out[-2000:]
    assert "--nproc_per_node 2" in out                              # devices: 2 read from the YAML
    exp = tmp_path / "exp"
    assert (tmp_path / "log").exists() and (exp / "cmd-args.log").exists()
    ck = sorted(os.listdir(exp / "checkpoints"))
    assert any("step=4" in c for c in ck), ck
    # relaunch with a larger budget: auto-resume continues from step 4
    out2 = _run(tmp_path, {"TRAIN_ITERS": "6"}, ["exp_manager.checkpoint_callback_params.every_n_train_steps=2"], 29782)
    steps2 = [int(m) for m in re.findall(r"step (\d+) \| reduced_train_loss", out2)]
    assert steps2 == [5, 6], out2[-2000:]
    assert "resumed from" in out2


def test_compile_dry_run_keeps_the_reference_behaviour(tmp_path):
    out = _run(tmp_path, {"COMPILE": "1"}, [], 29783)
    steps = [int(m) for m in re.findall(r"step (\d+) \| reduced_train_loss", out)]
    assert steps == list(range(1, 11))                              # COMPILE=1 -> 10 steps
    assert not (tmp_path / "exp" / "checkpoints").exists() or not os.listdir(tmp_path / "exp" / "checkpoints")
