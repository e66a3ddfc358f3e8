"""Model shards are read by ONE data-parallel rank and broadcast to the other replicas (nested dicts, mixed dtypes)."""
from common import *
from neuronx_distributed_training_b200.checkpoint import io as ckpt

rank, world = init(tp=1, pp=1)          # DP = world
work = os.environ["WORKDIR"]
torch.manual_seed(0)                    # same tensors on every rank (replicas)
model = {"w": torch.randn(33, 7), "h": {"b": torch.arange(10, dtype=torch.int64), "x": torch.randn(5).bfloat16()},
         "meta": {"note": "hello", "n": 3}, "lst": [torch.ones(2), 4.5]}
opt = {"step": 2, "exp_avg": torch.full((8,), float(rank))}
ckpt.save_checkpoint(work, "t.ckpt", model=model, optimizer=opt, user_content={"global_step": 2}, zero1_optimizer=True)
ckpt.finalize_checkpoint()
dist.barrier()

reads = []
orig = ckpt._load_object
def spy(path, *a, **k):
    reads.append(os.path.relpath(path, work))
    return orig(path, *a, **k)
ckpt._load_object = spy
m, o, uc = ckpt.load_checkpoint(work, "t.ckpt")
model_reads = [r for r in reads if "/model/" in r]
assert (len(model_reads) == 1) == (rank == 0), (rank, reads)            # only DP rank 0 touched the model shard
check("w", m["w"], model["w"], atol=0)
check("h.b", m["h"]["b"].float(), model["h"]["b"].float(), atol=0)
assert m["h"]["x"].dtype == torch.bfloat16 and torch.equal(m["h"]["x"].cpu(), model["h"]["x"])
assert m["meta"] == {"note": "hello", "n": 3} and m["lst"][1] == 4.5 and torch.equal(m["lst"][0].cpu(), torch.ones(2))
assert float(o["exp_avg"][0]) == float(rank)                              # optimizer shards stay per-rank
os.environ["NXDT_CKPT_LOAD_BROADCAST"] = "0"
reads.clear()
m2, _, _ = ckpt.load_checkpoint(work, "t.ckpt")
assert len([r for r in reads if "/model/" in r]) == 1                     # switch off: every rank reads
check("w again", m2["w"], model["w"], atol=0)
finish()
