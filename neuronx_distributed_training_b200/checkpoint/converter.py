"""Checkpoint conversion: HF / Megatron full-state checkpoints <-> TP/PP-sharded framework checkpoints.

Capability parity: NxD ``CheckpointConverterBase`` as driven by the reference's scripts (SURVEY §2.2 X16, §5.4;
reference examples/checkpoint_converter_scripts/*.py, docs/tutorials/checkpoint_conversion.rst:92-105): GQA
``qkv_linear`` / ``kv_size_multiplier`` / ``fuse_qkv`` aware, fused ``gate_up_proj`` with stride-2 sharding, global
layer numbering split over pipeline stages, xser or plain save, and the two user hooks
``pre_process_full_state_before_tp_conversion`` / ``post_process_full_state_after_tp_conversion``.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import re
from typing import Dict, List, Optional, Tuple

import torch

from . import io as ckpt_io
from ..parallel.layers import shard_master_weight, unshard_weights
from ..pipeline.partition import stage_layer_ranges

_LAYER_RE = re.compile(r"\.layers\.(\d+)\.")


# ---- file helpers ----------------------------------------------------------------------------------------
def load_full_state(path: str) -> Dict[str, torch.Tensor]:
    """HF directory (safetensors / bin shards), a single .pt/.bin/.safetensors file."""
    files: List[str] = []
    if os.path.isdir(path):
        for pat in ("*.safetensors", "pytorch_model*.bin", "*.pt"):
            files = sorted(glob.glob(os.path.join(path, pat)))
            if files:
                break
    else:
        files = [path]
    if not files:
        raise FileNotFoundError(f"no checkpoint files under {path}")
    state: Dict[str, torch.Tensor] = {}
    for f in files:
        if f.endswith(".safetensors"):
            from safetensors.torch import load_file
            state.update(load_file(f))
        else:
            obj = torch.load(f, map_location="cpu", weights_only=False)
            state.update(obj.get("state_dict", obj.get("model", obj)) if isinstance(obj, dict) else obj)
    return state


def save_full_state(state: Dict[str, torch.Tensor], out_dir: str, safetensors: bool = False) -> str:
    os.makedirs(out_dir, exist_ok=True)
    if safetensors:
        from safetensors.torch import save_file
        p = os.path.join(out_dir, "model.safetensors")
        save_file({k: v.contiguous() for k, v in state.items()}, p)
    else:
        p = os.path.join(out_dir, "pytorch_model.bin")
        torch.save(state, p)
    return p


# ---- converter --------------------------------------------------------------------------------------------
class CheckpointConverterBase:
    """Subclass and override the ``*_key_map`` / hooks for a new architecture."""

    # regex on the *sharded* (framework) parameter name -> (partition_dim, stride); None = replicated
    SHARD_RULES: List[Tuple[str, Optional[int], int]] = [
        (r"embed_tokens\.weight$", 0, 1), (r"lm_head\.weight$", 0, 1),
        (r"qkv_proj\.weight_(qkv|q|k|v)$", 0, 1), (r"qkv_proj\.bias_(qkv|q|k|v)$", 0, 1),
        (r"(q_proj|k_proj|v_proj)\.weight$", 0, 1), (r"o_proj\.weight$", 1, 1),
        (r"gate_up_proj\.(weight|bias)$", 0, 2), (r"down_proj\.weight$", 1, 1),
        (r"expert_mlps\.mlp_op\.gate_up_proj$", 2, 2), (r"expert_mlps\.mlp_op\.down_proj$", 1, 1),
        # Megatron names
        (r"word_embeddings\.weight$", 0, 1), (r"output_layer\.weight$", 0, 1),
        (r"query_key_value\.(weight|bias)$", 0, 1), (r"self_attention\.query\.(weight|bias)$", 0, 1),
        (r"key_value\.(weight|bias)$", 0, 1), (r"self_attention\.dense\.weight$", 1, 1),
        (r"dense_h_to_4h\.(weight|bias)$", 0, -1), (r"dense_4h_to_h\.weight$", 1, 1),
    ]

    def __init__(self, args: Optional[argparse.Namespace] = None, **kw):
        self.args = args or argparse.Namespace(**kw)
        for k, v in kw.items():
            setattr(self.args, k, v)
        a = self.args
        self.tp = int(getattr(a, "tp_size", 1))
        self.pp = int(getattr(a, "pp_size", 1))
        self.vpp = int(getattr(a, "virtual_pp_size", 1) or 1)
        self.kv_mult = int(getattr(a, "kv_size_multiplier", 1) or 1)
        self.qkv_linear = bool(getattr(a, "qkv_linear", True))
        self.fuse_qkv = bool(getattr(a, "fuse_qkv", True))
        self.model_style = str(getattr(a, "model_style", "hf"))
        self.config = self._load_config(getattr(a, "config", None))
        self.n_layers = int(getattr(a, "n_layers", 0) or self.config.get("num_hidden_layers", self.config.get("num_layers", 0)))
        self.glu = bool(getattr(a, "glu", True))

    @staticmethod
    def _load_config(path) -> Dict:
        if not path:
            return {}
        if isinstance(path, dict):
            return path
        with open(path) as f:
            return json.load(f)

    # ---- hooks (same names as the reference's base class) -------------------------------------------
    def pre_process_full_state_before_tp_conversion(self, state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return state

    def post_process_full_state_after_tp_conversion(self, state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return state

    # ---- HF <-> framework naming (full tensors) ------------------------------------------------------
    def _head_dims(self) -> Tuple[int, int, int]:
        c = self.config
        nh = int(c["num_attention_heads"])
        nkv = int(c.get("num_key_value_heads", nh) or nh)
        hd = int(c.get("head_dim") or (int(c["hidden_size"]) // nh))
        return nh, nkv, hd

    def replicate_kv(self, w: torch.Tensor) -> torch.Tensor:
        if self.kv_mult == 1:
            return w
        nh, nkv, hd = self._head_dims()
        return w.view(nkv, hd, -1).repeat_interleave(self.kv_mult, dim=0).reshape(nkv * self.kv_mult * hd, -1)

    def dereplicate_kv(self, w: torch.Tensor) -> torch.Tensor:
        if self.kv_mult == 1:
            return w
        nh, nkv, hd = self._head_dims()
        return w.view(nkv, self.kv_mult, hd, -1)[:, 0].reshape(nkv * hd, -1)

    def hf_to_framework_full(self, hf: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """HF names -> framework names with *full* (unsharded) tensors; q/k/v and gate/up stay separate here and are
        fused per TP rank in :meth:`shard_for_rank` (the fused layout interleaves per rank)."""
        if self.model_style != "hf":
            return dict(hf)
        out: Dict[str, torch.Tensor] = {}
        for k, v in hf.items():
            if k.endswith("rotary_emb.inv_freq"):
                continue
            out[k] = v
        return out

    # ---- sharding -------------------------------------------------------------------------------------------
    def _rule(self, name: str):
        for pat, dim, stride in self.SHARD_RULES:
            if re.search(pat, name):
                return dim, stride
        return None, 1

    def _layer_owner(self) -> Dict[int, int]:
        owner = {}
        if self.pp == 1:
            return {i: 0 for i in range(self.n_layers)}
        ranges = stage_layer_ranges(self.n_layers, self.pp, self.vpp, getattr(self.args, "pipeline_cuts", None))
        for r, chunks in enumerate(ranges):
            for lo, hi in chunks:
                for i in range(lo, hi):
                    owner[i] = r
        return owner

    def _pp_rank_of(self, name: str, owner: Dict[int, int]) -> int:
        m = _LAYER_RE.search(name)
        if m:
            return owner[int(m.group(1))]
        if "embed_tokens" in name or "embedding." in name:
            return 0
        return self.pp - 1          # final norm, lm_head / output_layer

    def shard_for_rank(self, full: Dict[str, torch.Tensor], tp_rank: int, pp_rank: int) -> Dict[str, torch.Tensor]:
        owner = self._layer_owner()
        out: Dict[str, torch.Tensor] = {}
        tp = self.tp
        done = set()
        for name, w in full.items():
            if name in done or self._pp_rank_of(name, owner) != pp_rank:
                continue
            if self.model_style == "hf":
                m = re.match(r"(.*self_attn)\.(q_proj|k_proj|v_proj)\.(weight|bias)$", name)
                if m and self.qkv_linear:
                    base, kind = m.group(1), m.group(3)
                    q, k, v = (full[f"{base}.{p}.{kind}"] for p in ("q_proj", "k_proj", "v_proj"))
                    done.update(f"{base}.{p}.{kind}" for p in ("q_proj", "k_proj", "v_proj"))
                    if kind == "bias":
                        q, k, v = q.unsqueeze(-1), k.unsqueeze(-1), v.unsqueeze(-1)
                    qs = shard_master_weight(q, 0, tp, tp_rank)
                    ks = shard_master_weight(self.replicate_kv(k), 0, tp, tp_rank)
                    vs = shard_master_weight(self.replicate_kv(v), 0, tp, tp_rank)
                    if kind == "bias":
                        qs, ks, vs = qs.squeeze(-1), ks.squeeze(-1), vs.squeeze(-1)
                    if self.fuse_qkv:
                        out[f"{base}.qkv_proj.{kind}_qkv"] = torch.cat([qs, ks, vs], 0)
                    else:
                        out[f"{base}.qkv_proj.{kind}_q"], out[f"{base}.qkv_proj.{kind}_k"], out[f"{base}.qkv_proj.{kind}_v"] = qs, ks, vs
                    continue
                m = re.match(r"(.*mlp)\.(gate_proj|up_proj)\.(weight|bias)$", name)        # bias: models with mlp_bias
                if m:
                    base, kind = m.group(1), m.group(3)
                    g, u = full[f"{base}.gate_proj.{kind}"], full[f"{base}.up_proj.{kind}"]
                    done.update((f"{base}.gate_proj.{kind}", f"{base}.up_proj.{kind}"))
                    out[f"{base}.gate_up_proj.{kind}"] = shard_master_weight(torch.cat([g, u], 0), 0, tp, tp_rank, stride=2)
                    continue
            dim, stride = self._rule(name)
            if stride == -1:
                stride = 2 if self.glu else 1
            out[name] = w if dim is None or tp == 1 else shard_master_weight(w, dim, tp, tp_rank, stride)
        return out

    def unshard(self, shards: List[List[Dict[str, torch.Tensor]]]) -> Dict[str, torch.Tensor]:
        """``shards[pp][tp]`` framework state dicts -> full HF (or Megatron) state."""
        full: Dict[str, torch.Tensor] = {}
        for pp_shards in shards:
            names = pp_shards[0].keys()
            for name in names:
                parts = [s[name] for s in pp_shards]
                if self.model_style == "hf":
                    m = re.match(r"(.*self_attn)\.qkv_proj\.(weight|bias)_(qkv|q|k|v)$", name)
                    if m:
                        base, kind, which = m.groups()
                        nh, nkv, hd = self._head_dims()
                        ql, kl = nh * hd // self.tp, nkv * self.kv_mult * hd // self.tp
                        if which == "qkv":
                            qs = [p[:ql] for p in parts]; ks = [p[ql:ql + kl] for p in parts]; vs = [p[ql + kl:] for p in parts]
                            full[f"{base}.q_proj.{kind}"] = torch.cat(qs, 0)
                            for nm, lst in (("k_proj", ks), ("v_proj", vs)):
                                w = torch.cat(lst, 0)
                                w = self.dereplicate_kv(w if kind == "weight" else w.unsqueeze(-1))
                                full[f"{base}.{nm}.{kind}"] = w if kind == "weight" else w.squeeze(-1)
                        else:
                            w = torch.cat(parts, 0)
                            if which in ("k", "v"):
                                w = self.dereplicate_kv(w if kind == "weight" else w.unsqueeze(-1))
                                w = w if kind == "weight" else w.squeeze(-1)
                            full[f"{base}.{which}_proj.{kind}"] = w
                        continue
                    m = re.match(r"(.*mlp)\.gate_up_proj\.(weight|bias)$", name)
                    if m:
                        gu = unshard_weights(parts, 0, stride=2)
                        g, u = gu.chunk(2, 0)
                        full[f"{m.group(1)}.gate_proj.{m.group(2)}"], full[f"{m.group(1)}.up_proj.{m.group(2)}"] = g, u
                        continue
                dim, stride = self._rule(name)
                if stride == -1:
                    stride = 2 if self.glu else 1
                full[name] = parts[0] if dim is None or self.tp == 1 else unshard_weights(parts, dim, stride)
        return full

    # ---- drivers ---------------------------------------------------------------------------------------------
    def convert_from_full_state(self, input_dir: str, output_dir: str, tag: str = "converted.ckpt", use_xser: bool = True):
        full = self.hf_to_framework_full(load_full_state(input_dir))
        full = self.pre_process_full_state_before_tp_conversion(full)
        root = os.path.join(output_dir, tag) if tag else output_dir
        os.makedirs(os.path.join(root, "model"), exist_ok=True)
        for pp_rank in range(self.pp):
            for tp_rank in range(self.tp):
                shard = self.shard_for_rank(full, tp_rank, pp_rank)
                shard = {k: v.clone().contiguous() for k, v in shard.items()}
                ckpt_io._save_object(shard, os.path.join(root, "model", ckpt_io.shard_file_name(0, tp_rank, pp_rank)), use_xser, False)
        with open(os.path.join(root, "checkpoint"), "w") as f:
            f.write((tag or os.path.basename(root)) + "\n")
        open(os.path.join(root, "done"), "w").close()
        return root

    def convert_to_full_state(self, input_dir: str, output_dir: str, safetensors: bool = False):
        """``input_dir`` is a ``*.ckpt`` directory (or its ``model/`` sub-directory)."""
        mdir = input_dir if os.path.basename(input_dir.rstrip("/")) == "model" else os.path.join(input_dir, "model")
        shards = [[ckpt_io._load_object(os.path.join(mdir, ckpt_io.shard_file_name(0, t, p))) for t in range(self.tp)]
                  for p in range(self.pp)]
        full = self.unshard(shards)
        full = self.post_process_full_state_after_tp_conversion(full)
        return save_full_state(full, output_dir, safetensors), full

    def run(self):
        a = self.args
        if getattr(a, "convert_from_full_state", False):
            return self.convert_from_full_state(a.input_dir, a.output_dir, getattr(a, "tag", "converted.ckpt"),
                                                bool(getattr(a, "save_xser", True)))
        if getattr(a, "convert_to_full_state", False):
            return self.convert_to_full_state(a.input_dir, a.output_dir, bool(getattr(a, "safetensors", False)))[0]
        raise ValueError("pass --convert_from_full_state or --convert_to_full_state")

    @staticmethod
    def get_arg_parser() -> argparse.ArgumentParser:
        p = argparse.ArgumentParser(description="HF / Megatron <-> sharded checkpoint converter")
        p.add_argument("--model_style", choices=["hf", "megatron"], default="hf")
        p.add_argument("--input_dir", required=True)
        p.add_argument("--output_dir", required=True)
        p.add_argument("--save_xser", type=lambda s: str(s).lower() in ("1", "true", "yes"), default=True)
        p.add_argument("--load_xser", type=lambda s: str(s).lower() in ("1", "true", "yes"), default=True)
        p.add_argument("--convert_from_full_state", action="store_true")
        p.add_argument("--convert_to_full_state", action="store_true")
        p.add_argument("--config", help="HF config.json of the model")
        p.add_argument("--tp_size", type=int, default=1)
        p.add_argument("--pp_size", type=int, default=1)
        p.add_argument("--virtual_pp_size", type=int, default=1)
        p.add_argument("--n_layers", type=int, default=0)
        p.add_argument("--kv_size_multiplier", type=int, default=1)
        p.add_argument("--qkv_linear", type=lambda s: str(s).lower() in ("1", "true", "yes"), default=True)
        p.add_argument("--fuse_qkv", type=lambda s: str(s).lower() in ("1", "true", "yes"), default=True)
        p.add_argument("--hf_model_name", default=None)
        p.add_argument("--hw_backend", default="b200", help="accepted for CLI compatibility (trn1/trn2 in the reference)")
        p.add_argument("--safetensors", action="store_true")
        p.add_argument("--tag", default="converted.ckpt")
        return p


class MixtralCheckpointConverter(CheckpointConverterBase):
    """HF Mixtral per-expert ``block_sparse_moe.experts.N.w{1,2,3}`` <-> stacked
    ``mlp.moe.expert_mlps.mlp_op.{gate_up,down}_proj`` and ``block_sparse_moe.gate`` <-> ``mlp.moe.router.linear_router``
    (reference hf_nxdt_mixtral_ckpt_converter.py:37-85)."""

    def pre_process_full_state_before_tp_conversion(self, state):
        out = {}
        E = int(self.config.get("num_local_experts", 8))
        layers = sorted({int(m.group(1)) for k in state for m in [_LAYER_RE.search(k)] if m and "block_sparse_moe" in k})
        for k, v in state.items():
            if "block_sparse_moe" not in k:
                out[k] = v
        # transformers >= 5 keeps (and may save) the experts fused: ``mlp.gate.weight`` [E,H],
        # ``mlp.experts.gate_up_proj`` [E, 2I, H] (gate rows first) and ``mlp.experts.down_proj`` [E, H, I]
        for k in [k for k in out if re.search(r"\.mlp\.(gate\.weight|experts\.(gate_up_proj|down_proj))$", k)]:
            v = out.pop(k)
            l = _LAYER_RE.search(k).group(1)
            if k.endswith("mlp.gate.weight"):
                out[f"model.layers.{l}.mlp.moe.router.linear_router.weight"] = v
            elif k.endswith("experts.gate_up_proj"):
                out[f"model.layers.{l}.mlp.moe.expert_mlps.mlp_op.gate_up_proj"] = v.transpose(1, 2).contiguous()    # [E,H,2I]
            else:
                out[f"model.layers.{l}.mlp.moe.expert_mlps.mlp_op.down_proj"] = v.transpose(1, 2).contiguous()       # [E,I,H]
        for l in layers:
            pre = f"model.layers.{l}.block_sparse_moe"
            out[f"model.layers.{l}.mlp.moe.router.linear_router.weight"] = state[f"{pre}.gate.weight"]
            w1 = torch.stack([state[f"{pre}.experts.{e}.w1.weight"].t() for e in range(E)])   # gate  [E,H,I]
            w3 = torch.stack([state[f"{pre}.experts.{e}.w3.weight"].t() for e in range(E)])   # up    [E,H,I]
            w2 = torch.stack([state[f"{pre}.experts.{e}.w2.weight"].t() for e in range(E)])   # down  [E,I,H]
            out[f"model.layers.{l}.mlp.moe.expert_mlps.mlp_op.gate_up_proj"] = torch.cat([w1, w3], dim=2)
            out[f"model.layers.{l}.mlp.moe.expert_mlps.mlp_op.down_proj"] = w2
        return out

    def post_process_full_state_after_tp_conversion(self, state):
        out = {}
        for k, v in state.items():
            m = re.match(r"model\.layers\.(\d+)\.mlp\.moe\.(.*)$", k)
            if not m:
                out[k] = v
                continue
            l, rest = m.groups()
            pre = f"model.layers.{l}.block_sparse_moe"
            if rest == "router.linear_router.weight":
                out[f"{pre}.gate.weight"] = v
            elif rest == "expert_mlps.mlp_op.gate_up_proj":
                g, u = v.chunk(2, dim=2)
                for e in range(v.shape[0]):
                    out[f"{pre}.experts.{e}.w1.weight"] = g[e].t().contiguous()
                    out[f"{pre}.experts.{e}.w3.weight"] = u[e].t().contiguous()
            elif rest == "expert_mlps.mlp_op.down_proj":
                for e in range(v.shape[0]):
                    out[f"{pre}.experts.{e}.w2.weight"] = v[e].t().contiguous()
        return out
