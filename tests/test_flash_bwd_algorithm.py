"""Tile-level emulation of ``flash_attn_bwd2_sm100.cu`` in PyTorch: same loop structure (one 128-key block per "CTA", query
blocks at or after it, every head of the GQA group), same formulas (exp2 domain, LSE / Delta statistics, window bounds).
It validates the algorithm the kernel implements against autograd on the CPU; the CUDA kernel itself is checked on the GPU by
``tests/test_kernels_gpu.py::test_flash_attention_forward_backward`` and ``tests/gpu_checks/check_flash_bwd.py``."""
import math

import pytest
import torch

from neuronx_distributed_training_b200 import ops

BM = BN = 128


def emulate(do, q, k, v, out, lse, window, scale):
    B, S, Hq, D = q.shape
    Hk = k.shape[2]
    rep = Hq // Hk
    nb = S // BM
    scale_log2 = scale * 1.4426950408889634
    delta = (do.float() * out.float()).sum(-1).permute(0, 2, 1)                 # [B, Hq, S]  (pre-pass kernel)
    dq = torch.zeros(B, S, Hq, D)
    dk = torch.zeros(B, S, Hk, D)
    dv = torch.zeros(B, S, Hk, D)
    for b in range(B):
        for kvh in range(Hk):
            for j in range(nb):                                                  # one CTA
                k0 = j * BN
                Kj, Vj = k[b, k0:k0 + BN, kvh].float(), v[b, k0:k0 + BN, kvh].float()
                i_hi = nb - 1
                if window > 0:
                    i_hi = min(i_hi, (k0 + BN - 1 + window - 1) // BM)
                dKj, dVj = torch.zeros(BN, D), torch.zeros(BN, D)
                for h in range(kvh * rep, (kvh + 1) * rep):
                    for i in range(j, i_hi + 1):
                        q0 = i * BM
                        Qi, dOi = q[b, q0:q0 + BM, h].float(), do[b, q0:q0 + BM, h].float()
                        lse2 = lse[b, h, q0:q0 + BM] * 1.4426950408889634
                        dl = delta[b, h, q0:q0 + BM]
                        Sij = Qi @ Kj.t()
                        dP = dOi @ Vj.t()
                        P = torch.exp2(Sij * scale_log2 - lse2[:, None])
                        qpos = torch.arange(q0, q0 + BM)[:, None]
                        kpos = torch.arange(k0, k0 + BN)[None, :]
                        mask = kpos > qpos
                        if window > 0:
                            mask = mask | (kpos <= qpos - window)
                        P = P.masked_fill(mask, 0.0)
                        P = P.masked_fill(~torch.isfinite(lse2)[:, None], 0.0)
                        dS = P * (dP - dl[:, None]) * scale
                        Pb, dSb = P.bfloat16().float(), dS.bfloat16().float()    # the kernel stages P / dS as bf16 tiles
                        dVj += Pb.t() @ dOi
                        dKj += dSb.t() @ Qi
                        dq[b, q0:q0 + BM, h] += dSb @ Kj
                dk[b, k0:k0 + BN, kvh] = dKj
                dv[b, k0:k0 + BN, kvh] = dVj
    return dq, dk, dv


@pytest.mark.parametrize("B,S,Hq,Hk,window", [(1, 256, 2, 1, 0), (2, 384, 4, 2, 0), (1, 512, 2, 2, 200), (1, 256, 1, 1, 128)])
def test_tiled_backward_matches_autograd(B, S, Hq, Hk, window):
    torch.manual_seed(0)
    q = torch.randn(B, S, Hq, 128).bfloat16().requires_grad_(True)
    k = torch.randn(B, S, Hk, 128).bfloat16().requires_grad_(True)
    v = torch.randn(B, S, Hk, 128).bfloat16().requires_grad_(True)
    do = torch.randn(B, S, Hq, 128).bfloat16()
    scale = 1.0 / math.sqrt(128)
    ref, lse = ops.attention_ref(q, k, v, causal=True, window=window or None, scale=scale, return_lse=True)
    ref.backward(do)
    dq, dk, dv = emulate(do, q.detach(), k.detach(), v.detach(), ref.detach(), lse.detach(), window, scale)
    for got, want in ((dq, q.grad), (dk, k.grad), (dv, v.grad)):
        err = ((got - want.float()).norm() / want.float().norm()).item()
        assert err < 2e-2, err
