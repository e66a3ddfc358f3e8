import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from neuronx_distributed_training_b200.parallel import state as ps


def init(tp=1, pp=1, cp=1, ep=1, vpp=1):
    ps.init_distributed("gloo")
    ps.initialize_model_parallel(tp, pp, cp, ep, vpp)
    torch.manual_seed(1234)
    return ps.get_rank(), ps.get_world_size()


def check(name, a, b, atol=1e-5, rtol=1e-4):
    ok = torch.allclose(a.float(), b.float(), atol=atol, rtol=rtol)
    if not ok:
        err = (a.float() - b.float()).abs().max().item()
        raise AssertionError(f"[rank {ps.get_rank()}] {name}: max abs err {err}")


def finish():
    dist.barrier()
    if ps.get_rank() == 0:
        print("DIST_TEST_OK")
    dist.destroy_process_group()
