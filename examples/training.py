"""``train(cfg)``: build strategy / trainer / exp-manager / data module / model module and fit.

Same selection rules as the reference entry point (reference examples/training.py:41-94):
``cfg.model_source`` in {hf, megatron}, presence of ``model_alignment_strategy``, and
``"mixtral" in cfg.name``.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from neuronx_distributed_training_b200.config import Config, ParallelDims
from neuronx_distributed_training_b200.parallel import state as ps
from neuronx_distributed_training_b200.trainer import Trainer
from neuronx_distributed_training_b200.utils.exp_manager import exp_manager


def build_data_module(cfg, trainer):
    source = cfg.get("model_source", "hf")
    if cfg.get("model_alignment_strategy", None) is not None or \
            (cfg.data.get("use_sft_style_data_module", False) and cfg.data.get("train_dir", None)):
        # the Megatron SFT config selects the HF-style (jsonl / arrow) data module with data.use_sft_style_data_module
        if cfg.get("model_alignment_strategy", None) is None:
            cfg["model_alignment_strategy"] = Config({"sft": Config({"packing": bool(cfg.data.get("packing", False))})})
        from neuronx_distributed_training_b200.data.alignment import ModelAlignmentDataModule
        return ModelAlignmentDataModule(cfg, trainer)
    if cfg.data.get("synthetic", False) or (source == "hf" and not cfg.data.get("train_dir", None)) \
            or (source == "megatron" and not cfg.data.get("data_prefix", None)):
        from neuronx_distributed_training_b200.data import SyntheticDataModule
        return SyntheticDataModule(cfg, trainer)
    if source == "megatron":
        from neuronx_distributed_training_b200.data.megatron import MegatronDataModule
        return MegatronDataModule(cfg, trainer)
    from neuronx_distributed_training_b200.data import HFDataModule
    return HFDataModule(cfg, trainer)


def build_model_module(cfg, trainer):
    source = cfg.get("model_source", "hf")
    if source == "megatron":
        from neuronx_distributed_training_b200.trainer.megatron_module import MegatronGPTModel
        return MegatronGPTModel(cfg, trainer)
    if "mixtral" in str(cfg.get("name", "")):
        from neuronx_distributed_training_b200.trainer.hf_mixtral_module import HFMixtralModule
        return HFMixtralModule(cfg, trainer)
    from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule
    return HFLlamaModule(cfg, trainer)


def train(cfg: Config):
    ps.init_distributed()
    dims = ParallelDims.from_config(cfg, ps.dist.get_world_size() if ps.dist.is_initialized() else 1)
    ps.initialize_model_parallel(dims.tp, dims.pp, dims.cp, dims.ep, dims.vpp)
    tr = cfg.trainer.to_dict()
    trainer = Trainer(**tr)
    em = exp_manager(trainer, cfg.get("exp_manager", None))
    data_module = build_data_module(cfg, trainer)
    model = build_model_module(cfg, trainer)
    trainer.fit(model, data_module, ckpt_path=em.resume_from_checkpoint)
    return trainer
