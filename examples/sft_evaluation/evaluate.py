"""Generation-based evaluation of fine-tuned weights (ROUGE / exact match).

    python evaluate.py --config ../conf/hf_llama3_8B_SFT_config.yaml --checkpoint <dir/*.ckpt | hf dir> \
        --data eval.jsonl --metric rouge --prompt_template "{{ input }}" --label_template "{{ output }}" \
        --max_new_tokens 64 --batch_size 8

Counterpart of the reference's examples/sft_evaluation/evaluate.py (:45-147): instead of separate inference runtimes the
fine-tuned (TP-sharded or HF-format) weights are loaded into the training model class (HF-Llama or HF-Mixtral, chosen
like the trainer does by ``"mixtral" in cfg.name``) and decoded in batches with a KV cache
(``utils/generation.py:generate_cached``).  Prompt / label templates are jinja2 (``{{ field }}``), as in the reference;
plain ``str.format`` templates (``{field}``) keep working.  Launch with torchrun when the checkpoint is TP-sharded."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from metrics import MetricFactory  # noqa: E402
from neuronx_distributed_training_b200.checkpoint import io as ckpt_io  # noqa: E402
from neuronx_distributed_training_b200.checkpoint.converter import CheckpointConverterBase, load_full_state  # noqa: E402
from neuronx_distributed_training_b200.config import ParallelDims, load_config, process_config  # noqa: E402
from neuronx_distributed_training_b200.parallel import state as ps  # noqa: E402
from neuronx_distributed_training_b200.utils.generation import generate_cached  # noqa: E402


def render(template: str, record: dict) -> str:
    """jinja2 when the template uses ``{{ }}`` / ``{% %}``, otherwise ``str.format``."""
    if "{{" in template or "{%" in template:
        import jinja2
        return jinja2.Environment(undefined=jinja2.StrictUndefined).from_string(template).render(**record)
    return template.format(**record)


def load_model(cfg, checkpoint: str):
    if "mixtral" in str(cfg.get("name", "")).lower():
        from neuronx_distributed_training_b200.trainer.hf_mixtral_module import HFMixtralModule as Module
    else:
        from neuronx_distributed_training_b200.trainer.hf_llama_module import HFLlamaModule as Module
    module = Module(cfg, None)
    module.setup("fit")
    if os.path.isdir(checkpoint) and os.path.isdir(os.path.join(checkpoint, "model")):
        sd = ckpt_io._load_object(os.path.join(checkpoint, "model", ckpt_io.shard_file_name(
            0, ps.get_tensor_model_parallel_rank(), ps.get_pipeline_model_parallel_rank())))
    else:   # HF-format weights: shard on the fly
        hf = module.hf_config
        conv = CheckpointConverterBase(tp_size=ps.get_tensor_model_parallel_size(), pp_size=1, model_style="hf",
                                       config={"num_attention_heads": hf.num_attention_heads, "num_key_value_heads": hf.num_key_value_heads,
                                               "hidden_size": hf.hidden_size, "head_dim": hf.head_dim, "num_hidden_layers": hf.num_hidden_layers},
                                       kv_size_multiplier=hf.kv_shared_group_size, fuse_qkv=hf.fuse_qkv, qkv_linear=hf.qkv_linear)
        sd = conv.shard_for_rank(conv.hf_to_framework_full(load_full_state(checkpoint)), ps.get_tensor_model_parallel_rank(), 0)
    module.load_state_dict(sd)
    return module


def evaluate(args):
    cfg = load_config(args.config, overrides=args.overrides)
    ps.init_distributed()
    world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
    cfg.distributed_strategy["pipeline_model_parallel_size"] = 1
    cfg.distributed_strategy["tensor_model_parallel_size"] = world
    process_config(cfg, world)
    ps.initialize_model_parallel(world)
    module = load_model(cfg, args.checkpoint)
    if args.tokenizer is not None:
        from transformers import AutoTokenizer
        tok = AutoTokenizer.from_pretrained(args.tokenizer)
    else:
        from transformers import AutoTokenizer
        tok = AutoTokenizer.from_pretrained(cfg.data.tokenizer.type)
    recs = [json.loads(l) for l in open(args.data) if l.strip()][: args.limit or None]
    preds, refs = [], []
    bs = max(1, int(args.batch_size))
    order = sorted(range(len(recs)), key=lambda i: len(render(args.prompt_template, recs[i])))     # similar lengths per batch
    prompts = {i: tok.encode(render(args.prompt_template, recs[i])) for i in order}
    out_text = {}
    for s0 in range(0, len(order), bs):
        idx = order[s0:s0 + bs]
        new = generate_cached(module.model, [prompts[i] for i in idx], args.max_new_tokens, tok.eos_token_id,
                              temperature=args.temperature, top_k=args.top_k, top_p=args.top_p,
                              pad_token_id=tok.pad_token_id if tok.pad_token_id is not None else 0)
        for i, ids in zip(idx, new):
            out_text[i] = tok.decode(ids, skip_special_tokens=True)
    for i, r in enumerate(recs):
        preds.append(out_text[i])
        refs.append(render(args.label_template, r))
    scores = MetricFactory.create(args.metric)(preds, refs)
    if ps.get_rank() == 0:
        print(json.dumps({"n": len(preds), **scores}))
        if args.output:
            with open(args.output, "w") as f:
                for p, r in zip(preds, refs):
                    f.write(json.dumps({"prediction": p, "reference": r}) + "\n")
    return scores


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--checkpoint", required=True)
    ap.add_argument("--data", required=True)
    ap.add_argument("--tokenizer", default=None)
    ap.add_argument("--metric", default="rouge")
    ap.add_argument("--prompt_template", default="{input}")
    ap.add_argument("--label_template", default="{output}")
    ap.add_argument("--max_new_tokens", type=int, default=64)
    ap.add_argument("--batch_size", type=int, default=8)
    ap.add_argument("--temperature", type=float, default=0.0)
    ap.add_argument("--top_k", type=int, default=0)
    ap.add_argument("--top_p", type=float, default=1.0)
    ap.add_argument("--limit", type=int, default=0)
    ap.add_argument("--output", default=None)
    ap.add_argument("overrides", nargs="*")
    evaluate(ap.parse_args())
