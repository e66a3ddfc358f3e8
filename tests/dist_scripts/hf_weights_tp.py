"""Weights of a `transformers` LlamaForCausalLM -> converter -> TP=2 shards -> this framework's model: the loss and the
gradient flowing into the embedding equal the transformers model's.  An independent check of the model math (RoPE
convention, GQA, SwiGLU, RMSNorm), of the converter's layouts (fused / split QKV, gate||up interleave) and of the
tensor-parallel layers at once."""
from common import *
import transformers
from neuronx_distributed_training_b200.checkpoint.converter import CheckpointConverterBase
from neuronx_distributed_training_b200.models.llama import LlamaConfig, LlamaForCausalLM

TP = int(os.environ.get("TP", "2"))
KVM = int(os.environ.get("KVM", "1"))        # kv_size_multiplier: TP=4 with 2 KV heads needs every KV head on two ranks
rank, world = init(tp=TP)
tp, tpr = ps.get_tensor_model_parallel_size(), ps.get_tensor_model_parallel_rank()
torch.manual_seed(3)                       # same HF model on every rank
hcfg = transformers.LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                                num_key_value_heads=2, max_position_embeddings=64, rope_theta=10000.0, rms_norm_eps=1e-5,
                                tie_word_embeddings=False, attention_bias=False, mlp_bias=False, attn_implementation="eager")
hf = transformers.LlamaForCausalLM(hcfg).float()
g = torch.Generator().manual_seed(5)
ids = torch.randint(0, 64, (2, 16), generator=g)
ref_loss = hf(input_ids=ids, labels=ids).loss
ref_loss.backward()
full = {k: v.detach().clone() for k, v in hf.state_dict().items()}
for fuse in (True, False):
    for sp in (False, True):
        conv = CheckpointConverterBase(tp_size=tp, pp_size=1, model_style="hf",
                                       config={"num_attention_heads": 4, "num_key_value_heads": 2, "hidden_size": 32, "head_dim": 8,
                                               "num_hidden_layers": 2},
                                       kv_size_multiplier=KVM, fuse_qkv=fuse, qkv_linear=True)
        shard = conv.shard_for_rank(conv.hf_to_framework_full(dict(full)), tpr, 0)
        m = LlamaForCausalLM(LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                                         num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=64, rope_theta=10000.0,
                                         rms_norm_eps=1e-5, torch_dtype="float32", use_flash_attention=False, fuse_qkv=fuse,
                                         kv_shared_group_size=KVM,
                                         sequence_parallel_enabled=sp, use_cpu_initialization=True))
        own = m.state_dict()
        missing = [k for k in own if k not in shard and not k.endswith("rotary_emb.inv_freq")]
        unexpected = [k for k in shard if k not in own]
        assert not missing and not unexpected, (missing[:4], unexpected[:4])
        m.load_state_dict(shard, strict=False)
        loss, _ = m(input_ids=ids, labels=ids)
        loss.backward()
        check(f"loss vs transformers fuse_qkv={fuse} sp={sp}", loss.detach(), ref_loss.detach(), atol=2e-5)
        from neuronx_distributed_training_b200.parallel.layers import shard_master_weight
        check(f"embedding grad fuse_qkv={fuse} sp={sp}", m.model.embed_tokens.weight.grad,
              shard_master_weight(hf.model.embed_tokens.weight.grad, 0, tp, tpr), atol=2e-5)
if KVM > 1:
    finish()
    sys.exit(0)
# RoPE scaling variants against transformers (forward loss, TP=2 + SP, fused QKV): Llama-3.1 style, linear, dynamic NTK
for rs in ({"rope_type": "llama3", "factor": 4.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0, "original_max_position_embeddings": 8},
           {"rope_type": "linear", "factor": 2.0}):
    torch.manual_seed(3)
    hc = transformers.LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                                  num_key_value_heads=2, max_position_embeddings=64, rope_theta=10000.0, rms_norm_eps=1e-5,
                                  tie_word_embeddings=False, rope_scaling=dict(rs), attn_implementation="eager")
    hm = transformers.LlamaForCausalLM(hc).float()
    with torch.no_grad():
        want = hm(input_ids=ids, labels=ids).loss
    conv = CheckpointConverterBase(tp_size=tp, pp_size=1, model_style="hf", kv_size_multiplier=1, fuse_qkv=True, qkv_linear=True,
                                   config={"num_attention_heads": 4, "num_key_value_heads": 2, "hidden_size": 32, "head_dim": 8,
                                           "num_hidden_layers": 2})
    shard = conv.shard_for_rank(conv.hf_to_framework_full({k: v.detach().clone() for k, v in hm.state_dict().items()}), tpr, 0)
    m = LlamaForCausalLM(LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                                     num_key_value_heads=2, max_position_embeddings=64, rope_theta=10000.0, rms_norm_eps=1e-5,
                                     rope_scaling=dict(rs), torch_dtype="float32", use_flash_attention=False, fuse_qkv=True,
                                     sequence_parallel_enabled=True, use_cpu_initialization=True))
    m.load_state_dict(shard, strict=False)
    with torch.no_grad():
        got, _ = m(input_ids=ids, labels=ids)
    check(f"rope_scaling {rs['rope_type']} loss vs transformers", got, want, atol=2e-5)
# attention / MLP biases (fused and split QKV)
for fuse in (True, False):
    torch.manual_seed(3)
    hc = transformers.LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                                  num_key_value_heads=2, max_position_embeddings=64, rope_theta=10000.0, rms_norm_eps=1e-5,
                                  tie_word_embeddings=False, attention_bias=True, mlp_bias=True, attn_implementation="eager")
    hm = transformers.LlamaForCausalLM(hc).float()
    with torch.no_grad():
        for n_, p_ in hm.named_parameters():
            if n_.endswith("bias"):
                p_.normal_(0, 0.1)                      # transformers initialises biases to zero
        want = hm(input_ids=ids, labels=ids).loss
    conv = CheckpointConverterBase(tp_size=tp, pp_size=1, model_style="hf", kv_size_multiplier=1, fuse_qkv=fuse, qkv_linear=True,
                                   config={"num_attention_heads": 4, "num_key_value_heads": 2, "hidden_size": 32, "head_dim": 8,
                                           "num_hidden_layers": 2})
    shard = conv.shard_for_rank(conv.hf_to_framework_full({k: v.detach().clone() for k, v in hm.state_dict().items()}), tpr, 0)
    m = LlamaForCausalLM(LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                                     num_key_value_heads=2, max_position_embeddings=64, rope_theta=10000.0, rms_norm_eps=1e-5,
                                     attention_bias=True, mlp_bias=True, torch_dtype="float32", use_flash_attention=False,
                                     fuse_qkv=fuse, sequence_parallel_enabled=True, use_cpu_initialization=True))
    own = m.state_dict()
    missing = [k for k in own if k not in shard and not k.endswith("rotary_emb.inv_freq")]
    unexpected = [k for k in shard if k not in own]
    assert not missing and not unexpected, ("bias model", missing[:4], unexpected[:4])
    m.load_state_dict(shard, strict=False)
    with torch.no_grad():
        got, _ = m(input_ids=ids, labels=ids)
    check(f"biased model loss vs transformers fuse_qkv={fuse}", got, want, atol=2e-5)
# sliding-window attention against transformers' Mistral (same parameter names as Llama)
torch.manual_seed(3)
mc = transformers.MistralConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                                num_key_value_heads=2, max_position_embeddings=64, rope_theta=10000.0, rms_norm_eps=1e-5,
                                sliding_window=5, tie_word_embeddings=False, attn_implementation="eager", head_dim=8)
mm_ = transformers.MistralForCausalLM(mc).float()
with torch.no_grad():
    want = mm_(input_ids=ids, labels=ids).loss
conv = CheckpointConverterBase(tp_size=tp, pp_size=1, model_style="hf", kv_size_multiplier=1, fuse_qkv=True, qkv_linear=True,
                               config={"num_attention_heads": 4, "num_key_value_heads": 2, "hidden_size": 32, "head_dim": 8,
                                       "num_hidden_layers": 2})
shard = conv.shard_for_rank(conv.hf_to_framework_full({k: v.detach().clone() for k, v in mm_.state_dict().items()}), tpr, 0)
m = LlamaForCausalLM(LlamaConfig(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                                 num_key_value_heads=2, max_position_embeddings=64, rope_theta=10000.0, rms_norm_eps=1e-5,
                                 sliding_window=5, torch_dtype="float32", use_flash_attention=False, fuse_qkv=True,
                                 sequence_parallel_enabled=True, use_cpu_initialization=True))
m.load_state_dict(shard, strict=False)
with torch.no_grad():
    got, _ = m(input_ids=ids, labels=ids)
check("sliding-window loss vs transformers Mistral", got, want, atol=2e-5)
finish()
