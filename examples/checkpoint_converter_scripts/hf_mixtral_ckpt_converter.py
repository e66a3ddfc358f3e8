"""CLI: HF Mixtral <-> sharded framework checkpoint (stacks / unstacks the per-expert w1/w2/w3 weights and maps the
router); extra flags ``--ep_size`` and ``--nxdt_yaml_config`` as in the reference script."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from neuronx_distributed_training_b200.checkpoint.converter import MixtralCheckpointConverter  # noqa: E402

if __name__ == "__main__":
    p = MixtralCheckpointConverter.get_arg_parser()
    p.add_argument("--ep_size", type=int, default=1)
    p.add_argument("--nxdt_yaml_config", default=None)
    args = p.parse_args()
    if args.nxdt_yaml_config:
        from neuronx_distributed_training_b200.config import load_config
        y = load_config(args.nxdt_yaml_config)
        ds = y.distributed_strategy
        args.tp_size = int(ds.get("tensor_model_parallel_size", args.tp_size))
        args.pp_size = int(ds.get("pipeline_model_parallel_size", args.pp_size))
        args.kv_size_multiplier = int(ds.get("kv_replicator", args.kv_size_multiplier) or 1)
    print(MixtralCheckpointConverter(args).run())
