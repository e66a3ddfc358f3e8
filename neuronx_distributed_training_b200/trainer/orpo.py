"""Odds-Ratio Preference Optimisation (no reference model).

Capability parity: reference lightning_modules/model/base_orpo.py:14-45:
``loss = NLL(chosen) - beta * log sigmoid(log_odds)`` with length-averaged sequence log-probs and
``log_odds = (logp_c - logp_r) - (log1p(-exp(logp_c)) - log1p(-exp(logp_r)))``.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .dpo import concatenated_inputs, sequence_logps


def orpo_loss(logp_c: torch.Tensor, logp_r: torch.Tensor, beta: float):
    log_odds = (logp_c - logp_r) - (torch.log1p(-torch.exp(logp_c).clamp(max=1 - 1e-7)) -
                                    torch.log1p(-torch.exp(logp_r).clamp(max=1 - 1e-7)))
    ratio = F.logsigmoid(log_odds)
    nll = -logp_c.mean()
    loss = nll - beta * ratio.mean()
    return loss, (beta * logp_c).detach(), (beta * logp_r).detach()


class ORPOBaseModel:
    @staticmethod
    def model_fwd_calc_loss(module, batch: Dict[str, torch.Tensor]):
        from .dpo import _reject_context_parallel
        _reject_context_parallel(module, "ORPO")
        ocfg = module.cfg.model_alignment_strategy.orpo
        beta = float(ocfg.get("beta", 0.1))
        logp_c, logp_r = sequence_logps(module.model, concatenated_inputs(batch), average=True)
        loss, cr, rr = orpo_loss(logp_c, logp_r, beta)
        module._logged["chosen_rewards"] = cr.mean().detach()      # read back only when metrics are logged
        module._logged["rejected_rewards"] = rr.mean().detach()
        return loss, None
