"""Entry point: ``training_orchestrator.py --config-path=conf --config-name=<stem> key=value ...``

CLI-compatible with the reference orchestrator (reference examples/training_orchestrator.py:140-145)
without hydra: ``--config-path``/``--config-name`` plus hydra-style overrides.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from neuronx_distributed_training_b200.config import load_config, process_config


def parse_args(argv):
    ap = argparse.ArgumentParser(add_help=True)
    ap.add_argument("--config-path", default="conf")
    ap.add_argument("--config-name", default="hf_llama3_8B_config")
    args, overrides = ap.parse_known_args(argv)
    bad = [o for o in overrides if "=" not in o and not o.startswith("~")]
    if bad:
        ap.error(f"unrecognised arguments (expected key=value overrides): {bad}")
    return args, overrides


def main(argv=None):
    args, overrides = parse_args(sys.argv[1:] if argv is None else argv)
    cfg_dir = args.config_path if os.path.isabs(args.config_path) else \
        os.path.join(os.path.dirname(os.path.abspath(__file__)), args.config_path)
    cfg = load_config(cfg_dir, args.config_name, [o for o in overrides if not o.endswith("=None") or True])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    process_config(cfg, world)
    from training import train
    train(cfg)


if __name__ == "__main__":
    main()
