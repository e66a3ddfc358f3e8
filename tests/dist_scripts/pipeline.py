"""1F1B / interleaved pipeline: loss and gradients equal the un-pipelined model."""
from common import *
from neuronx_distributed_training_b200.models.llama import LlamaConfig, LlamaForCausalLM
from neuronx_distributed_training_b200.parallel.model import ParallelModel
from neuronx_distributed_training_b200.pipeline.partition import stage_layer_ranges

PP = int(os.environ.get("PP", "2")); VPP = int(os.environ.get("VPP", "1")); TP = int(os.environ.get("TP", "1"))
rank, world = init(tp=TP, pp=PP, vpp=VPP)
ppr = ps.get_pipeline_model_parallel_rank()
L, NMB, MBS, S = 4 * VPP, int(os.environ.get("NMB", "4")), 1, 16
cfg = LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=L, num_attention_heads=4,
                  num_key_value_heads=2, torch_dtype="float32", use_flash_attention=False,
                  sequence_parallel_enabled=TP > 1, use_cpu_initialization=True)
torch.manual_seed(42)
full = LlamaForCausalLM(cfg)            # identical on every rank (cpu init from the same seed)
ranges = stage_layer_ranges(L, PP, VPP)[ppr]
chunks = []
for c, (lo, hi) in enumerate(ranges):
    first, last = ppr == 0 and c == 0, ppr == PP - 1 and c == VPP - 1
    ch = LlamaForCausalLM(cfg, layer_range=(lo, hi), with_embedding=first, with_head=last)
    sd = {k: v for k, v in full.state_dict().items() if k in ch.state_dict()}
    ch.load_state_dict(sd, strict=True)
    chunks.append(ch)
pm = ParallelModel(chunks, NMB, MBS, cfg.hidden_size, sequence_parallel=TP > 1, dtype=torch.float32)
g = torch.Generator().manual_seed(9)
ids = torch.randint(0, 64, (NMB * MBS, S), generator=g)
batch = {"input_ids": ids, "labels": ids.clone()}
loss = pm.run_train(**batch)
dist.all_reduce(loss, group=ps.get_pipeline_model_parallel_group())
# reference: microbatch loop on the full model
ref_loss = 0.0
for i in range(NMB):
    l, _ = full(input_ids=ids[i:i + 1], labels=ids[i:i + 1])
    (l / NMB).backward()
    ref_loss += l.item() / NMB
check("pp loss", loss, torch.tensor(ref_loss), atol=1e-5)
ref_grads = dict((n, p.grad) for n, p in full.named_parameters())
for n, p in pm.local_named_parameters():
    assert p.grad is not None, n
    check(f"grad {n}", p.grad, ref_grads[n], atol=2e-5)
ev = pm.run_eval(**batch)
dist.all_reduce(ev, group=ps.get_pipeline_model_parallel_group())
check("pp eval loss", ev, torch.tensor(ref_loss), atol=1e-5)
finish()
