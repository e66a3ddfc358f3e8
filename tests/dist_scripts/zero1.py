"""ZeRO-1 flat AdamW over DP (with grad clipping) vs torch.optim.AdamW on the full batch."""
from common import *
from neuronx_distributed_training_b200.optim import DistributedOptimizer

rank, world = init()
dp, dpr = ps.get_data_parallel_size(), ps.get_data_parallel_rank()
torch.manual_seed(0)
def make():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(12, 33), torch.nn.Tanh(), torch.nn.Linear(33, 5))
for zero1 in (True, False):
    m, ref = make(), make()
    no_decay = [p for n, p in m.named_parameters() if n.endswith("bias")]
    decay = [p for n, p in m.named_parameters() if not n.endswith("bias")]
    opt = DistributedOptimizer([{"params": decay, "weight_decay": 0.1}, {"params": no_decay, "weight_decay": 0.0}],
                               lr=1e-2, betas=(0.9, 0.95), zero1=zero1, master_weights=True, fp32_grad_acc=True, max_grad_norm=0.5,
                               dp_group=ps.get_data_parallel_group())
    rdecay = [p for n, p in ref.named_parameters() if not n.endswith("bias")]
    rno = [p for n, p in ref.named_parameters() if n.endswith("bias")]
    ropt = torch.optim.AdamW([{"params": rdecay, "weight_decay": 0.1}, {"params": rno, "weight_decay": 0.0}], lr=1e-2, betas=(0.9, 0.95))
    g = torch.Generator().manual_seed(3)
    for step in range(4):
        X = torch.randn(4 * dp, 12, generator=g); Y = torch.randn(4 * dp, 5, generator=g)
        opt.zero_grad()
        xl, yl = X.chunk(dp)[dpr], Y.chunk(dp)[dpr]
        ((m(xl) - yl) ** 2).mean().backward()
        opt.step()
        ropt.zero_grad()
        ((ref(X) - Y) ** 2).mean().backward()
        gn = torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.5)
        ropt.step()
        check(f"grad norm zero1={zero1} step {step}", opt.grad_norm, gn, atol=1e-5)
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        check(f"param {n} zero1={zero1}", p, q, atol=2e-5)
    sd = opt.state_dict()
    assert sd["exp_avg"].numel() == opt.total // (dp if zero1 else 1)
finish()
