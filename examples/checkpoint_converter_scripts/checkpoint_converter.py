"""CLI: HF / Megatron full-state checkpoint <-> TP/PP-sharded framework checkpoint.

Same flag set as the reference's converter (docs/tutorials/checkpoint_conversion.rst:92-105):
    python checkpoint_converter.py --model_style hf --input_dir <hf_dir> --output_dir <out> --convert_from_full_state \
        --config <config.json> --tp_size 8 --pp_size 1 --n_layers 32 --kv_size_multiplier 1 --qkv_linear True --fuse_qkv True
    python checkpoint_converter.py --model_style hf --input_dir <ckpt_dir> --output_dir <hf_out> --convert_to_full_state ...
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from neuronx_distributed_training_b200.checkpoint.converter import CheckpointConverterBase  # noqa: E402


class CheckpointConverterLlama(CheckpointConverterBase):
    pass


if __name__ == "__main__":
    args = CheckpointConverterLlama.get_arg_parser().parse_args()
    print(CheckpointConverterLlama(args).run())
