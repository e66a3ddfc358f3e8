import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "sft_evaluation"))
from metrics import rouge_n, rouge_l, MetricFactory
from neuronx_distributed_training_b200.models.llama import LlamaConfig, LlamaForCausalLM
from neuronx_distributed_training_b200.utils.generation import generate
from neuronx_distributed_training_b200.utils.perf_estimate import llama_flops_per_seq, calculate_mfu


def test_rouge_and_em():
    assert rouge_n("the cat sat", "the cat sat", 1) == 1.0
    assert abs(rouge_n("the cat", "the cat sat", 1) - 0.8) < 1e-9
    assert abs(rouge_l("a b c d", "a c d") - 2 * 0.75 * 1.0 / 1.75) < 1e-9
    assert MetricFactory.create("em")(["Yes", "no"], ["yes", "maybe"])["exact_match"] == 0.5
    assert 0 < MetricFactory.create("rouge")(["the quick fox"], ["the quick brown fox"])["rougeL"] < 1


def test_greedy_generation_is_consistent_with_forward():
    torch.manual_seed(0)
    m = LlamaForCausalLM(LlamaConfig(vocab_size=50, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                                     num_key_value_heads=2, torch_dtype="float32", use_flash_attention=False))
    ids = torch.randint(0, 50, (2, 5))
    out = generate(m, ids, max_new_tokens=4)
    assert out.shape == (2, 9) and torch.equal(out[:, :5], ids)
    _, logits = m(input_ids=out[:, :-1], labels=None)
    assert torch.equal(logits[:, 4:].argmax(-1), out[:, 5:])


def test_llama3_8b_flops_model():
    f = llama_flops_per_seq(4096, 32, 8192, 128256, 14336, 32, 8, 128)
    assert 4.5e14 < f < 5.5e14                      # ~5.0e14 FLOPs / 8k-token sequence
    assert abs(calculate_mfu(1.0, f, 1, 1000.0) - f / 1e15) < 1e-12


def test_top_k_and_top_p_filtering():
    import torch
    from neuronx_distributed_training_b200.utils.generation import filter_logits
    logits = torch.tensor([[2.0, 1.0, 0.5, -1.0, 3.0]])
    f = filter_logits(logits, top_k=2)
    assert torch.isfinite(f).sum() == 2 and torch.isfinite(f[0, [0, 4]]).all()
    p = torch.softmax(logits, -1)
    f = filter_logits(logits, top_p=float(p[0, 4]) + 1e-4)                   # nucleus = the best token + the next one
    assert torch.isfinite(f).sum() == 2 and torch.isfinite(f[0, [0, 4]]).all()
    f = filter_logits(logits, top_p=0.01)                                    # the most likely token always survives
    assert torch.isfinite(f).sum() == 1 and torch.isfinite(f[0, 4])
    assert torch.equal(filter_logits(logits), logits)


def test_kv_cached_batched_generation_matches_uncached():
    """Variable-length prompts, left-padded and decoded with the KV cache, give exactly the tokens of the prefix-re-running
    decoder run on every prompt alone (Llama with GQA + sliding window, and Mixtral)."""
    from neuronx_distributed_training_b200.models.mixtral import MixtralConfig, MixtralForCausalLM
    from neuronx_distributed_training_b200.utils.generation import generate_cached, generate_uncached, supports_kv_cache
    torch.manual_seed(1)
    models = [
        LlamaForCausalLM(LlamaConfig(vocab_size=61, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                                     num_key_value_heads=2, torch_dtype="float32", use_flash_attention=False, sliding_window=6)),
        MixtralForCausalLM(MixtralConfig(vocab_size=61, hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=4,
                                         num_key_value_heads=2, torch_dtype="float32", use_flash_attention=False,
                                         num_local_experts=4, num_experts_per_tok=2)),
    ]
    prompts = [[3, 9, 27, 4, 8, 15, 16, 23, 42], [5], [7, 7, 7, 1]]
    for m in models:
        assert supports_kv_cache(m)
        new = generate_cached(m, prompts, max_new_tokens=6)
        for p, n in zip(prompts, new):
            ref = generate_uncached(m, torch.tensor([p]), max_new_tokens=6)[0, len(p):].tolist()
            assert n == ref, (type(m).__name__, p, n, ref)
    # eos stops a row early and is not part of the returned text
    m = models[0]
    first = generate_cached(m, prompts, max_new_tokens=6)
    eos = first[0][2]
    cut = generate_cached(m, prompts, max_new_tokens=6, eos_token_id=eos)
    assert cut[0] == first[0][:first[0].index(eos)]
