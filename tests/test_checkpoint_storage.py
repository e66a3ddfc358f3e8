"""Object-store checkpoint paths through a directory-backed store registered for the ``s3`` scheme."""
import os

import pytest
import torch

from neuronx_distributed_training_b200.checkpoint import io as ckpt
from neuronx_distributed_training_b200.checkpoint import storage
from neuronx_distributed_training_b200.parallel import state as ps


@pytest.fixture()
def fake_s3(tmp_path, monkeypatch):
    ps.initialize_model_parallel(1, 1, 1)
    root = tmp_path / "buckets"
    storage.register_object_store("s3", lambda: storage.DirectoryStore(str(root)))
    monkeypatch.setenv("NXDT_CKPT_STAGING", str(tmp_path / "staging"))
    yield root
    ckpt.finalize_checkpoint()
    storage.register_object_store("s3", storage.S3Store)
    ps.destroy_model_parallel()


def _state():
    torch.manual_seed(0)
    return ({"w": torch.randn(4, 5), "nested": {"b": torch.arange(6).float()}},
            {"step": 3, "exp_avg": torch.randn(20)}, {"global_step": 3, "note": "x"})


@pytest.mark.parametrize("async_save,use_xser", [(False, True), (True, True), (False, False)])
def test_save_load_roundtrip_through_object_store(fake_s3, async_save, use_xser):
    m, o, uc = _state()
    url = "s3://my-bucket/run1/checkpoints"
    out = ckpt.save_checkpoint(url, "step_3", m, o, uc, use_xser=use_xser, async_save=async_save)
    assert out == url + "/step_3"
    ckpt.finalize_checkpoint()
    assert ckpt.has_checkpoint(url, "step_3") and not ckpt.has_checkpoint(url, "step_4")
    keys = storage.get_store(url).list(url + "/step_3")
    assert any(k.endswith("/done") for k in keys) and any("/model/dp_rank_00_tp_rank_00_pp_rank_00.pt" in k for k in keys)
    assert any("/optim/" in k for k in keys) and any(k.endswith("user_content.pt") for k in keys)
    m2, o2, uc2 = ckpt.load_checkpoint(url, "step_3")
    assert torch.equal(m2["w"], m["w"]) and torch.equal(m2["nested"]["b"], m["nested"]["b"])
    assert torch.equal(o2["exp_avg"], o["exp_avg"]) and o2["step"] == 3 and uc2["note"] == "x"
    # staging holds no leftover shard files
    left = [f for _, _, fs in os.walk(os.environ["NXDT_CKPT_STAGING"]) for f in fs]
    assert left == [], left
    ckpt.remove_checkpoint(url + "/step_3", async_remove=False)
    assert not ckpt.has_checkpoint(url, "step_3")


def test_extra_text_files_reach_the_object_store(fake_s3):
    """user_content["_extra_files"] (the PEFT adapter_config.json) is stored as plain files beside user_content.pt"""
    m, o, uc = _state()
    uc[ckpt.EXTRA_FILES_KEY] = {"adapter_config.json": '{"peft_type": "LORA"}'}
    url = "s3://my-bucket/run2/checkpoints"
    ckpt.save_checkpoint(url, "step_1", m, o, uc)
    ckpt.finalize_checkpoint()
    keys = storage.get_store(url).list(url + "/step_1")
    assert any(k.endswith("/adapter_config.json") for k in keys), keys
    assert not any(k.endswith(".extra_files") for k in keys)
    _, _, uc2 = ckpt.load_checkpoint(url, "step_1", model=True, optimizer=True)
    assert ckpt.EXTRA_FILES_KEY not in uc2 and uc2["note"] == "x"


def test_incomplete_remote_checkpoint_is_rejected(fake_s3):
    m, o, uc = _state()
    url = "s3://b/ck"
    ckpt.save_checkpoint(url, "t", m, o, uc)
    os.remove(fake_s3 / "b" / "ck" / "t" / "done")
    with pytest.raises(FileNotFoundError):
        ckpt.load_checkpoint(url, "t")


def test_s3_without_boto3_gives_actionable_error(monkeypatch):
    storage.register_object_store("s3", storage.S3Store)
    try:
        import boto3  # noqa: F401
        pytest.skip("boto3 present")
    except ImportError:
        pass
    with pytest.raises(RuntimeError, match="boto3"):
        storage.get_store("s3://bucket/x")


class _FakeS3Client:
    """In-memory stand-in for ``boto3.client('s3')`` with the calls S3Store makes; list_objects_v2 pages by 3 keys so that
    the continuation-token loop is exercised, delete_objects rejects more than 1000 keys like the real service."""
    PAGE = 3

    def __init__(self):
        self.objects = {}                                   # (bucket, key) -> bytes
        self.delete_calls = []

    def upload_file(self, local_path, bucket, key):
        with open(local_path, "rb") as f:
            self.objects[(bucket, key)] = f.read()

    def download_file(self, bucket, key, local_path):
        with open(local_path, "wb") as f:
            f.write(self.objects[(bucket, key)])

    def list_objects_v2(self, Bucket, Prefix, ContinuationToken=None, MaxKeys=None):
        keys = sorted(k for (b, k) in self.objects if b == Bucket and k.startswith(Prefix))
        start = int(ContinuationToken or 0)
        page = self.PAGE if MaxKeys is None else min(self.PAGE, MaxKeys)
        chunk = keys[start:start + page]
        more = start + page < len(keys) and MaxKeys is None
        r = {"Contents": [{"Key": k} for k in chunk], "IsTruncated": more}
        if more:
            r["NextContinuationToken"] = str(start + page)
        return r

    def delete_objects(self, Bucket, Delete):
        assert len(Delete["Objects"]) <= 1000
        self.delete_calls.append(len(Delete["Objects"]))
        for o in Delete["Objects"]:
            del self.objects[(Bucket, o["Key"])]


@pytest.fixture()
def fake_boto3(monkeypatch, tmp_path):
    import sys, types
    client = _FakeS3Client()
    mod = types.ModuleType("boto3")
    mod.client = lambda name: client
    monkeypatch.setitem(sys.modules, "boto3", mod)
    monkeypatch.setenv("NXDT_CKPT_STAGING", str(tmp_path / "staging"))
    ps.initialize_model_parallel(1, 1, 1)
    storage.register_object_store("s3", storage.S3Store)
    yield client
    ckpt.finalize_checkpoint()
    ps.destroy_model_parallel()


def test_s3store_calls_against_a_fake_boto3_client(fake_boto3, tmp_path):
    """The real S3Store class (not the directory emulation): upload / download, paginated listing, exact-key `exists`
    (a key that is only a PREFIX of another object does not exist), batched prefix removal."""
    s3 = storage.S3Store()
    src = tmp_path / "blob.bin"
    for i in range(8):
        src.write_bytes(bytes([i]) * (i + 1))
        s3.upload_file(str(src), f"s3://bkt/run/ckpt/part{i}.pt")
    s3.upload_file(str(src), "s3://bkt/run/ckpt2/other.pt")
    listed = s3.list("s3://bkt/run/ckpt")
    assert listed == [f"s3://bkt/run/ckpt/part{i}.pt" for i in range(8)]          # 3 pages, other prefixes excluded
    assert s3.exists("s3://bkt/run/ckpt/part3.pt") and not s3.exists("s3://bkt/run/ckpt/part") \
        and not s3.exists("s3://bkt/run/ckpt/missing.pt")
    dst = tmp_path / "back.bin"
    s3.download_file("s3://bkt/run/ckpt/part5.pt", str(dst))
    assert dst.read_bytes() == bytes([5]) * 6
    s3.remove_prefix("s3://bkt/run/ckpt")
    assert s3.list("s3://bkt/run/ckpt") == [] and s3.exists("s3://bkt/run/ckpt2/other.pt")
    assert fake_boto3.delete_calls == [8]


def test_checkpoint_roundtrip_through_s3store_with_fake_boto3(fake_boto3):
    """save -> has -> load -> remove of a checkpoint with the real S3Store driving the (fake) boto3 client."""
    m, o, uc = _state()
    url = "s3://bkt/exp/checkpoints"
    ckpt.save_checkpoint(url, "step_3", m, o, uc, use_xser=True, async_save=False)
    ckpt.finalize_checkpoint()
    assert ckpt.has_checkpoint(url, "step_3") and not ckpt.has_checkpoint(url, "step_4")
    assert any(k.endswith("/done") for (_, k) in fake_boto3.objects)
    m2, o2, uc2 = ckpt.load_checkpoint(url, "step_3")
    assert torch.equal(m2["w"], m["w"]) and torch.equal(m2["nested"]["b"], m["nested"]["b"])
    assert torch.equal(o2["exp_avg"], o["exp_avg"]) and o2["step"] == 3 and uc2["note"] == "x"
    ckpt.remove_checkpoint(url + "/step_3", async_remove=False)
    assert not ckpt.has_checkpoint(url, "step_3") and not fake_boto3.objects
