"""CLI for Megatron-family checkpoints.

1. Full-state <-> sharded conversion (``--model_style megatron``), same flags as ``checkpoint_converter.py``.
2. ``--local_to_global_layers``: NeMo-Megatron (NNM) per-rank model checkpoints -> this framework's sharded layout
   (reference nnm_model_ckpt_to_nxdt_model_ckpt_converter.py: ``--tp --pp --n_layers --nnm_ckpt_path --nxdt_ckpt_path
   [--enable_parallel_processing --num_parallel_processes]``; the same flag names are accepted here).

    python megatron_ckpt_converter.py --local_to_global_layers --tp 8 --pp 4 --n_layers 32 \
        --nnm_ckpt_path /ckpts/nnm/step_1000 --nxdt_ckpt_path /ckpts/converted/step_1000
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from neuronx_distributed_training_b200.checkpoint.converter import CheckpointConverterBase  # noqa: E402
from neuronx_distributed_training_b200.checkpoint import nnm  # noqa: E402


def _bool(s):
    return str(s).lower() in ("1", "true", "yes")


def nnm_main(argv):
    p = argparse.ArgumentParser(description="NNM per-rank model checkpoint -> sharded framework checkpoint")
    p.add_argument("--local_to_global_layers", action="store_true")
    p.add_argument("--tp", type=int, required=True, help="tensor-parallel degree of the NNM run")
    p.add_argument("--pp", type=int, required=True, help="pipeline-parallel degree of the NNM run")
    p.add_argument("--n_layers", type=int, required=True, help="total number of transformer layers")
    p.add_argument("--nnm_ckpt_path", required=True, help="directory holding tp_rank_XX_pp_rank_YYY/model_optim_rng.ckpt")
    p.add_argument("--nxdt_ckpt_path", required=True, help="output checkpoint directory (a _vN sibling is used if not empty)")
    p.add_argument("--pipeline_cuts", nargs="*", default=None, help="uneven stage boundaries, as in the training YAML")
    p.add_argument("--enable_parallel_processing", type=_bool, default=False)
    p.add_argument("--num_parallel_processes", type=int, default=max(1, (os.cpu_count() or 8) - 5))
    p.add_argument("--save_xser", type=_bool, default=True, help="tensor-per-file layout (default) or one torch file per shard")
    a = p.parse_args(argv)
    return nnm.convert(a.nnm_ckpt_path, a.nxdt_ckpt_path, a.tp, a.pp, a.n_layers,
                       a.num_parallel_processes if a.enable_parallel_processing else 0, a.pipeline_cuts, a.save_xser)


if __name__ == "__main__":
    if "--local_to_global_layers" in sys.argv[1:]:
        print(nnm_main(sys.argv[1:]))
    else:
        p = CheckpointConverterBase.get_arg_parser()
        p.set_defaults(model_style="megatron")
        p.add_argument("--glu", type=_bool, default=True)
        print(CheckpointConverterBase(p.parse_args()).run())
