"""Adafactor with ZeRO-1 sharding (factored statistics assembled from the shards) == the unsharded reference optimizer."""
from common import *
from neuronx_distributed_training_b200.optim import DistributedOptimizer
from transformers.optimization import Adafactor

rank, world = init()
dp, dpr = ps.get_data_parallel_size(), ps.get_data_parallel_rank()
def make():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(12, 33), torch.nn.Tanh(), torch.nn.Linear(33, 5))
m, ref = make(), make()
opt = DistributedOptimizer([{"params": list(m.parameters())}], lr=1e-2, weight_decay=0.01, algorithm="adafactor", zero1=True,
                           master_weights=True, fp32_grad_acc=True, max_grad_norm=None, eps=(1e-30, 1e-3),
                           dp_group=ps.get_data_parallel_group(), extra_defaults={"beta1": 0.9, "scale_parameter": True})
assert opt.zero1 and opt.shard_size < opt.total
ropt = Adafactor(list(ref.parameters()), lr=1e-2, eps=(1e-30, 1e-3), beta1=0.9, weight_decay=0.01, scale_parameter=True,
                 relative_step=False, warmup_init=False)
g = torch.Generator().manual_seed(3)
for step in range(5):
    X = torch.randn(4 * dp, 12, generator=g); Y = torch.randn(4 * dp, 5, generator=g)
    opt.zero_grad()
    xl, yl = X.chunk(dp)[dpr], Y.chunk(dp)[dpr]
    ((m(xl) - yl) ** 2).mean().backward()
    opt.step()
    ropt.zero_grad()
    ((ref(X) - Y) ** 2).mean().backward()
    ropt.step()
for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
    check(f"param {n}", p, q, atol=2e-5)
# checkpoint round trip of the factored state
sd = opt.state_dict()
opt2_m = make()
opt2 = DistributedOptimizer([{"params": list(opt2_m.parameters())}], lr=1e-2, weight_decay=0.01, algorithm="adafactor", zero1=True,
                            master_weights=True, fp32_grad_acc=True, max_grad_norm=None, eps=(1e-30, 1e-3),
                            dp_group=ps.get_data_parallel_group(), extra_defaults={"beta1": 0.9, "scale_parameter": True})
opt2.load_state_dict(sd)
check("factored state", opt2.af_rc, opt.af_rc, atol=0)
finish()
