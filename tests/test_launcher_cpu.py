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


def _source_setup(env, tmp_path):
    """source examples/train_setup.sh in a clean bash with `env` and return the variables it exports"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "examples", "train_setup.sh")
    e = {"PATH": os.environ["PATH"], "HOME": str(tmp_path)}
    e.update(env)
    out = subprocess.run(["bash", "-c", f"cd {tmp_path} && source {script} && "
                          "echo NTASKS=$NTASKS NODEID=$NODEID MASTER_ADDR=$MASTER_ADDR MASTER_PORT=$MASTER_PORT "
                          "LOG_PATH=$LOG_PATH VERSION=$NEMO_EXPM_VERSION"],
                         env=e, capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    return dict(kv.split("=", 1) for kv in out.stdout.split())


def test_train_setup_slurm_branch(tmp_path):
    """SLURM: one task per node, rank = SLURM_NODEID, master = first host of the node list (through `scontrol`), logs and
    experiment version keyed by the job id (reference train_setup.sh behaviour)."""
    fake = tmp_path / "bin"
    fake.mkdir()
    sc = fake / "scontrol"
    sc.write_text("#!/bin/bash\n# scontrol show hostnames <list>\necho node-a17\necho node-a18\n")
    sc.chmod(0o755)
    v = _source_setup({"PATH": f"{fake}:{os.environ['PATH']}", "SLURM_NTASKS": "2", "SLURM_NODEID": "1", "SLURM_JOB_ID": "4711",
                       "SLURM_JOB_NODELIST": "node-a[17-18]"}, tmp_path)
    assert v["NTASKS"] == "2" and v["NODEID"] == "1" and v["MASTER_ADDR"] == "node-a17" and v["MASTER_PORT"] == "41000"
    assert v["LOG_PATH"] == "logs/4711/0/1" and v["VERSION"] == "4711"
    assert (tmp_path / "logs" / "4711" / "0" / "1").is_dir()


def test_train_setup_openmpi_branch(tmp_path):
    """OpenMPI / mpirun-style launchers: world size and rank from OMPI_*, master address parsed out of the HNP URI."""
    v = _source_setup({"OMPI_COMM_WORLD_SIZE": "4", "OMPI_COMM_WORLD_RANK": "3",
                       "OMPI_MCA_orte_hnp_uri": "1234567.0;tcp://10.1.2.3,172.17.0.1:40123"}, tmp_path)
    assert v["NTASKS"] == "4" and v["NODEID"] == "3" and v["MASTER_ADDR"] == "10.1.2.3"
    assert v["LOG_PATH"] == "logs/ompi/3" and v["VERSION"] == "version_0"


def test_train_setup_single_node_defaults(tmp_path):
    v = _source_setup({}, tmp_path)
    assert v["NTASKS"] == "1" and v["NODEID"] == "0" and v["MASTER_ADDR"] == "127.0.0.1" and v["LOG_PATH"] == "logs/local"
