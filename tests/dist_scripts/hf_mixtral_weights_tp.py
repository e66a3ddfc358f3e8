"""Weights of a `transformers` MixtralForCausalLM -> Mixtral converter -> TP=2 shards -> this framework's Mixtral: loss
and embedding gradient equal the transformers model's.  Independent check of the routing semantics (softmax -> top-2 ->
renormalise), of the expert weight layout (fused `mlp.experts.gate_up_proj` of transformers >= 5) and of the TP-sharded
experts.  The auxiliary loss is switched off for this comparison: transformers 5.x feeds already-softmaxed router
probabilities into `load_balancing_loss_func`, which applies softmax again, whereas the reference (and this framework)
evaluate it on the raw router logits; the values agree near uniform routing but the gradients do not."""
from common import *
import transformers
from neuronx_distributed_training_b200.checkpoint.converter import MixtralCheckpointConverter
from neuronx_distributed_training_b200.models.mixtral import MixtralConfig, MixtralForCausalLM
from neuronx_distributed_training_b200.parallel.layers import shard_master_weight

TP = int(os.environ.get("TP", "2"))
rank, world = init(tp=TP)
tp, tpr = ps.get_tensor_model_parallel_size(), ps.get_tensor_model_parallel_rank()
torch.manual_seed(3)
hcfg = transformers.MixtralConfig(vocab_size=64, hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=4,
                                  num_key_value_heads=2, max_position_embeddings=64, num_local_experts=4, num_experts_per_tok=2,
                                  router_aux_loss_coef=0.0, output_router_logits=True, sliding_window=None, rope_theta=10000.0,
                                  rms_norm_eps=1e-5, attn_implementation="eager")
hf = transformers.MixtralForCausalLM(hcfg).float()
g = torch.Generator().manual_seed(5)
ids = torch.randint(0, 64, (2, 16), generator=g)
out = hf(input_ids=ids, labels=ids)
out.loss.backward()
full = {k: v.detach().clone() for k, v in hf.state_dict().items()}
conv = MixtralCheckpointConverter(tp_size=tp, pp_size=1, model_style="hf", kv_size_multiplier=1, fuse_qkv=True, qkv_linear=True,
                                  config={"num_attention_heads": 4, "num_key_value_heads": 2, "hidden_size": 32, "head_dim": 8,
                                          "num_hidden_layers": 2, "num_local_experts": 4})
fw = conv.pre_process_full_state_before_tp_conversion(conv.hf_to_framework_full(dict(full)))
shard = conv.shard_for_rank(fw, tpr, 0)
for sp in (False, True):
    cfg = MixtralConfig(vocab_size=64, hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=4,
                        num_key_value_heads=2, max_position_embeddings=64, rope_theta=10000.0, rms_norm_eps=1e-5,
                        torch_dtype="float32", use_flash_attention=False, fuse_qkv=True, sequence_parallel_enabled=sp,
                        use_cpu_initialization=True, num_local_experts=4, num_experts_per_tok=2, router_aux_loss_coef=0.0)
    m = MixtralForCausalLM(cfg)
    own = m.state_dict()
    missing = [k for k in own if k not in shard and not k.endswith("rotary_emb.inv_freq")]
    unexpected = [k for k in shard if k not in own]
    assert not missing and not unexpected, (missing[:4], unexpected[:4])
    m.load_state_dict(shard, strict=False)
    loss, _ = m(input_ids=ids, labels=ids)
    loss.backward()
    check(f"loss vs transformers sp={sp}", loss.detach(), out.loss.detach(), atol=3e-5)
    check(f"embedding grad sp={sp}", m.model.embed_tokens.weight.grad,
          shard_master_weight(hf.model.embed_tokens.weight.grad, 0, tp, tpr), atol=3e-5)
finish()
