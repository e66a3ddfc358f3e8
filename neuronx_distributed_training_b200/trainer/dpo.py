"""Direct Preference Optimisation.

Capability parity: reference lightning_modules/model/base_dpo.py:14-142 - before training the (still un-updated)
policy is run over the whole training set to pre-compute the reference chosen / rejected log-probs, which are
attached to the dataset as columns; the training forward concatenates chosen and rejected on the batch dimension;
log-probs come straight from the vocab-parallel logits; loss = -log sigmoid(beta * (pi_logratio - ref_logratio));
``chosen_rewards`` / ``rejected_rewards`` are logged.
"""
from __future__ import annotations

import logging
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from ..parallel import state as ps
from ..parallel.loss import from_parallel_logits_to_logprobs

log = logging.getLogger("nxdt")


def concatenated_inputs(batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """chosen || rejected along the batch dimension (both already padded to the same length)."""
    return {"input_ids": torch.cat([batch["chosen_input_ids"], batch["rejected_input_ids"]], 0),
            "attention_mask": torch.cat([batch["chosen_attention_mask"], batch["rejected_attention_mask"]], 0),
            "labels": torch.cat([batch["chosen_labels"], batch["rejected_labels"]], 0)}


def sequence_logps(model, cat: Dict[str, torch.Tensor], average: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """sum (or mean) over answer tokens of log p(token); returns (chosen, rejected)."""
    _, logits = model(input_ids=cat["input_ids"], labels=None)          # [2B, S, V/TP]
    labels = cat["labels"]
    pad = torch.full((labels.shape[0], 1), -100, dtype=labels.dtype, device=labels.device)
    shifted = torch.cat([labels[:, 1:], pad], dim=1)                    # next-token alignment
    lp = from_parallel_logits_to_logprobs(logits, shifted)              # [2B, S], 0 where ignored
    mask = (shifted != -100).to(lp.dtype)
    tot = (lp * mask).sum(-1)
    if average:
        tot = tot / mask.sum(-1).clamp(min=1)
    b = labels.shape[0] // 2
    return tot[:b], tot[b:]


def dpo_loss(pi_c, pi_r, ref_c, ref_r, beta: float, loss_type: str = "sigmoid", label_smoothing: float = 0.0):
    logits = (pi_c - pi_r) - (ref_c - ref_r)
    if loss_type == "sigmoid":
        loss = -F.logsigmoid(beta * logits) * (1 - label_smoothing) - F.logsigmoid(-beta * logits) * label_smoothing
    elif loss_type == "hinge":
        loss = torch.relu(1 - beta * logits)
    elif loss_type == "ipo":
        loss = (logits - 1 / (2 * beta)) ** 2
    else:
        raise ValueError(f"unknown dpo loss_type {loss_type}")
    chosen_rewards = beta * (pi_c - ref_c).detach()
    rejected_rewards = beta * (pi_r - ref_r).detach()
    return loss.mean(), chosen_rewards, rejected_rewards


def _reject_context_parallel(module, what: str) -> None:
    """Sequence log-probs sum over the whole sequence; with the sequence split over CP ranks the per-rank partial sums
    would have to be reduced before the sigmoid.  The reference runs its alignment recipes with CP = 1 only."""
    if getattr(module, "dims", None) is not None and module.dims.cp > 1:
        raise ValueError(f"{what} does not support context_parallel_size > 1; set distributed_strategy.context_parallel_size=1")


class DPOBaseModel:
    @staticmethod
    @torch.no_grad()
    def on_train_start(module, trainer):
        """Pre-compute reference log-probs with the frozen initial policy and add them as dataset columns."""
        dcfg = module.cfg.model_alignment_strategy.dpo
        _reject_context_parallel(module, "DPO")
        if not bool(dcfg.get("precompute_ref_log_probs", True)):
            return
        dm = trainer.datamodule
        ds = dm.train_dataset()
        if ds is None or "reference_chosen_logps" in getattr(ds, "extra", {}):
            return
        module.model.eval()
        n = len(ds)
        ref_c = torch.zeros(n, dtype=torch.float32)
        ref_r = torch.zeros(n, dtype=torch.float32)
        mbs = module.micro_batch_size
        dp, dpr = ps.get_data_parallel_size(), ps.get_data_parallel_rank()
        idxs = list(range(dpr, n, dp))
        for s in range(0, len(idxs), mbs):
            chunk = idxs[s:s + mbs]
            items = [ds[i] for i in chunk]
            batch = {k: torch.stack([it[k] for it in items]).to(module.device) for k in items[0] if k.endswith(("_input_ids", "_attention_mask", "_labels"))}
            with module.autocast_ctx():
                c, r = sequence_logps(module.model, concatenated_inputs(batch))
            ref_c[chunk] = c.float().cpu()
            ref_r[chunk] = r.float().cpu()
        if dp > 1:
            both = torch.stack([ref_c, ref_r]).to(module.device)
            torch.distributed.all_reduce(both, group=ps.get_data_parallel_group())
            ref_c, ref_r = both[0].cpu(), both[1].cpu()
        ds.add_column("reference_chosen_logps", ref_c.tolist())
        ds.add_column("reference_rejected_logps", ref_r.tolist())
        module.model.train()
        log.info("DPO: pre-computed reference log-probs for %d samples", n)

    @staticmethod
    def model_fwd_calc_loss(module, batch: Dict[str, torch.Tensor]):
        dcfg = module.cfg.model_alignment_strategy.dpo
        _reject_context_parallel(module, "DPO")
        beta = float(dcfg.get("kl_beta", dcfg.get("beta", 0.1)))
        pi_c, pi_r = sequence_logps(module.model, concatenated_inputs(batch))
        ref_c = batch["reference_chosen_logps"].to(pi_c.dtype)
        ref_r = batch["reference_rejected_logps"].to(pi_c.dtype)
        loss, cr, rr = dpo_loss(pi_c, pi_r, ref_c, ref_r, beta, str(dcfg.get("loss_type", "sigmoid")))
        module._logged["chosen_rewards"] = cr.mean().detach()      # read back only when metrics are logged
        module._logged["rejected_rewards"] = rr.mean().detach()
        module._logged["reward_margin"] = (cr - rr).mean().detach()
        module._logged["reward_accuracy"] = (cr > rr).float().mean().detach()
        return loss, None
