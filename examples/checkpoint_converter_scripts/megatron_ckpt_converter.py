"""CLI: Megatron/NeMo-style full-state GPT checkpoint <-> sharded framework checkpoint (``--model_style megatron``).

A NeMo-Megatron checkpoint stores per-PP-stage *local* layer indices; ``--local_to_global_layers`` rewrites them to the
global numbering this framework uses (reference nnm_model_ckpt_to_nxdt_model_ckpt_converter.py:76-152)."""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
from neuronx_distributed_training_b200.checkpoint.converter import CheckpointConverterBase  # noqa: E402
from neuronx_distributed_training_b200.pipeline.partition import stage_layer_ranges  # noqa: E402


def local_to_global(state, pp_rank, pp_size, n_layers):
    lo, _ = stage_layer_ranges(n_layers, pp_size)[pp_rank][0]
    out = {}
    for k, v in state.items():
        out[re.sub(r"\.layers\.(\d+)\.", lambda m: f".layers.{int(m.group(1)) + lo}.", k)] = v
    return out


if __name__ == "__main__":
    p = CheckpointConverterBase.get_arg_parser()
    p.set_defaults(model_style="megatron")
    p.add_argument("--glu", type=lambda s: str(s).lower() in ("1", "true", "yes"), default=True)
    args = p.parse_args()
    print(CheckpointConverterBase(args).run())
