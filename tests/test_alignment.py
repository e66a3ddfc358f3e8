import json
import math

import torch
import torch.nn.functional as F

from neuronx_distributed_training_b200.config import from_dict
from neuronx_distributed_training_b200.data.alignment import (tokenize_sft, tokenize_preference, ConcatDataset, PaddedDataset,
                                                              PaddedDPODataset, ModelAlignmentDataModule, IGNORE_INDEX)
from neuronx_distributed_training_b200.models.llama import LlamaConfig, LlamaForCausalLM
from neuronx_distributed_training_b200.models import lora
from neuronx_distributed_training_b200.trainer.dpo import dpo_loss, sequence_logps, concatenated_inputs
from neuronx_distributed_training_b200.trainer.orpo import orpo_loss


class CharTok:
    bos_token_id, eos_token_id, pad_token_id = 1, 2, 0

    def encode(self, text, add_special_tokens=False):
        return [3 + (ord(c) % 60) for c in text]


def test_sft_tokenisation_masks_prompt():
    s = tokenize_sft({"input": "abc", "output": "de"}, CharTok(), 32)
    assert s["input_ids"][0] == 1 and s["input_ids"][-1] == 2 and len(s["input_ids"]) == 1 + 3 + 2 + 1
    assert s["labels"][:4] == [IGNORE_INDEX] * 4 and s["labels"][4:] == s["input_ids"][4:]


def test_packing_keeps_each_column():
    tok = CharTok()
    samples = [tokenize_sft({"input": "ab", "output": "c" * n}, tok, 16) for n in (2, 3, 9, 1)]
    ds = ConcatDataset(samples, 16, 0)
    for i in range(len(ds)):
        row = ds[i]
        assert row["input_ids"].shape == (16,)
        # labels are either -100 or equal to the token at the same position (NOT a copy of input_ids everywhere)
        lab, ids = row["labels"], row["input_ids"]
        assert ((lab == IGNORE_INDEX) | (lab == ids)).all()
        assert (lab == IGNORE_INDEX).any() and (lab != IGNORE_INDEX).any()
        assert ((row["attention_mask"] == 0) | (row["attention_mask"] == 1)).all()
    assert len(ds) < len(samples)                      # something got packed together
    pd = PaddedDataset(samples, 16, 0)
    assert pd[0]["attention_mask"].sum() == len(samples[0]["input_ids"])


def test_preference_tokenisation_and_padding():
    tok = CharTok()
    rec = {"prompt": "p" * 30, "chosen": "c" * 10, "rejected": "r" * 4}
    s = tokenize_preference(rec, tok, seq_length=24, max_prompt_length=8, truncation_mode="keep_end")
    assert len(s["prompt_input_ids"]) == 8 and len(s["chosen_input_ids"]) <= 24
    assert s["chosen_labels"][:8] == [IGNORE_INDEX] * 8
    ds = PaddedDPODataset([s], 24, 8, 0)
    ds.add_column("reference_chosen_logps", [-1.5])
    it = ds[0]
    assert it["chosen_input_ids"].shape == (24,) and it["prompt_input_ids"].shape == (8,)
    assert float(it["reference_chosen_logps"]) == -1.5


def test_dpo_and_orpo_losses():
    pc, pr = torch.tensor([-1.0, -2.0]), torch.tensor([-3.0, -2.5])
    rc, rr = torch.tensor([-1.5, -2.0]), torch.tensor([-2.5, -2.0])
    loss, cr, rj = dpo_loss(pc, pr, rc, rr, beta=0.1)
    ref = -F.logsigmoid(0.1 * ((pc - pr) - (rc - rr))).mean()
    assert torch.allclose(loss, ref) and torch.allclose(cr, 0.1 * (pc - rc))
    lo, _, _ = orpo_loss(torch.tensor([-0.5]), torch.tensor([-1.5]), 0.2)
    odds = (-0.5 + 1.5) - (math.log1p(-math.exp(-0.5)) - math.log1p(-math.exp(-1.5)))
    assert abs(float(lo) - (0.5 - 0.2 * math.log(1 / (1 + math.exp(-odds))))) < 1e-5


def _tiny(**kw):
    cfg = LlamaConfig(vocab_size=80, hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, torch_dtype="float32", use_flash_attention=False, **kw)
    torch.manual_seed(0)
    return LlamaForCausalLM(cfg)


def test_sequence_logps_matches_manual():
    m = _tiny()
    ids = torch.randint(3, 80, (2, 12))
    labels = ids.clone(); labels[:, :5] = -100
    batch = {"chosen_input_ids": ids[:1], "rejected_input_ids": ids[1:], "chosen_attention_mask": torch.ones(1, 12, dtype=torch.long),
             "rejected_attention_mask": torch.ones(1, 12, dtype=torch.long), "chosen_labels": labels[:1], "rejected_labels": labels[1:]}
    c, r = sequence_logps(m, concatenated_inputs(batch))
    _, logits = m(input_ids=ids, labels=None)
    lp = F.log_softmax(logits.float(), -1)[:, :-1].gather(-1, ids[:, 1:].unsqueeze(-1)).squeeze(-1)
    mask = (labels[:, 1:] != -100).float()
    ref = (lp * mask).sum(-1)
    assert torch.allclose(torch.stack([c[0], r[0]]), ref, atol=1e-4)


def test_lora_injection_and_merge():
    m = _tiny()
    ids = torch.randint(3, 80, (2, 10))
    cfg = lora.LoraConfig(lora_rank=4, lora_alpha=8, lora_dropout=0.0, target_modules=["qkv_proj", "o_proj", "gate_up_proj", "down_proj"])
    wrapped = lora.inject_lora(m, cfg)
    assert len(wrapped) == 8
    trainable = [n for n, p in m.named_parameters() if p.requires_grad]
    assert trainable and all(".adapter.lora_" in n for n in trainable)
    base_loss, _ = m(input_ids=ids, labels=ids)
    # B starts at zero => adapters are a no-op at init
    for mod in m.modules():
        if isinstance(mod, lora.LoraLinear):
            torch.nn.init.normal_(mod.adapter.lora_B, std=0.05)
    l1, _ = m(input_ids=ids, labels=ids)
    assert abs(float(l1) - float(base_loss)) > 1e-6
    l1.backward()
    assert all(p.grad is not None for n, p in m.named_parameters() if p.requires_grad)
    sd = lora.merged_state_dict(m)
    assert not any("adapter" in k or "base_layer" in k for k in sd)
    plain = _tiny()
    plain.load_state_dict(sd)
    l2, _ = plain(input_ids=ids, labels=ids)
    assert abs(float(l1) - float(l2)) < 1e-4
    assert set(lora.lora_state_dict(m)) == {n for n in m.state_dict() if ".adapter.lora_" in n}


def test_alignment_datamodule_sft_and_dpo(tmp_path):
    p = tmp_path / "sft.jsonl"
    p.write_text("\n".join(json.dumps({"input": "q" * (i % 5 + 1), "output": "a" * (i % 7 + 1)}) for i in range(20)))
    cfg = from_dict({"seed": 1, "data": {"micro_batch_size": 1, "global_batch_size": 2, "train_dir": str(p), "seq_length": 32},
                     "model": {"encoder_seq_length": 32}, "model_alignment_strategy": {"sft": {"packing": True}}})
    dm = ModelAlignmentDataModule(cfg, tokenizer=CharTok())
    dm.setup()
    b = next(iter(dm.train_dataloader()))
    assert b["input_ids"].shape == (2, 32) and b["labels"].shape == (2, 32)
    q = tmp_path / "dpo.jsonl"
    q.write_text("\n".join(json.dumps({"prompt": "p" * 4, "chosen": "c" * 5, "rejected": "r" * 3}) for _ in range(8)))
    cfg = from_dict({"seed": 1, "data": {"micro_batch_size": 1, "global_batch_size": 2, "train_dir": str(q), "seq_length": 16},
                     "model": {"encoder_seq_length": 16},
                     "model_alignment_strategy": {"dpo": {"kl_beta": 0.01, "max_prompt_length": 8, "truncation_mode": "keep_start"}}})
    dm = ModelAlignmentDataModule(cfg, tokenizer=CharTok())
    dm.setup()
    assert dm.mode == "dpo"
    b = next(iter(dm.train_dataloader()))
    assert b["chosen_input_ids"].shape == (2, 16) and b["prompt_input_ids"].shape == (2, 8)
