"""Object-store back ends for checkpoint directories (``s3://bucket/prefix`` paths).

Capability parity: the reference's checkpoint API accepts S3 URLs for save / load / keep-last-K removal
(docs/general/features.rst:57-66).  Design here: every rank writes its shard files to a local
staging directory with the normal (native, streamed) writer and then uploads exactly the files it produced; the
``done`` marker is uploaded last by rank 0, so a reader never sees a half-written checkpoint.  Loading downloads
only the files a rank needs.  Back ends are registered per URL scheme so other stores (or a test double) plug in
without touching the checkpoint code; the S3 one needs ``boto3`` and fails with a clear message without it.
"""
from __future__ import annotations

import hashlib
import os
import shutil
from typing import Callable, Dict, List, Optional, Tuple


def split_url(url: str) -> Tuple[str, str, str]:
    """``s3://bucket/a/b`` -> ("s3", "bucket", "a/b")."""
    scheme, rest = url.split("://", 1)
    bucket, _, key = rest.partition("/")
    return scheme, bucket, key.strip("/")


def is_remote(path: str) -> bool:
    return "://" in str(path) and not str(path).startswith("file://")


def join_url(url: str, *parts: str) -> str:
    return "/".join([url.rstrip("/")] + [p.strip("/") for p in parts if p])


class ObjectStore:
    """Minimal interface the checkpoint code needs."""

    def upload_file(self, local_path: str, url: str) -> None:
        raise NotImplementedError

    def download_file(self, url: str, local_path: str) -> None:
        raise NotImplementedError

    def list(self, url_prefix: str) -> List[str]:
        """All object URLs below ``url_prefix`` (recursive)."""
        raise NotImplementedError

    def exists(self, url: str) -> bool:
        raise NotImplementedError

    def remove_prefix(self, url_prefix: str) -> None:
        raise NotImplementedError

    # ---- helpers built on the primitives ----
    def upload_tree(self, local_root: str, url_root: str, only: Optional[Callable[[str], bool]] = None) -> int:
        n = 0
        for dirpath, _, files in os.walk(local_root):
            for f in files:
                lp = os.path.join(dirpath, f)
                rel = os.path.relpath(lp, local_root).replace(os.sep, "/")
                if only is not None and not only(rel):
                    continue
                self.upload_file(lp, join_url(url_root, rel))
                n += 1
        return n

    def download_prefix(self, url_prefix: str, local_root: str, base_url: str) -> int:
        n = 0
        for u in self.list(url_prefix):
            rel = u[len(base_url.rstrip("/")) + 1:]
            lp = os.path.join(local_root, *rel.split("/"))
            os.makedirs(os.path.dirname(lp), exist_ok=True)
            self.download_file(u, lp)
            n += 1
        return n


class S3Store(ObjectStore):
    def __init__(self):
        try:
            import boto3  # noqa: F401
        except ImportError as e:  # pragma: no cover - boto3 is not in the offline image
            raise RuntimeError("s3:// checkpoint paths need the 'boto3' package (not installed in this environment); "
                               "install it, mount the bucket (e.g. mountpoint-s3) and pass the mount path, or register "
                               "another back end with checkpoint.storage.register_object_store('s3', factory)") from e
        import boto3
        self._s3 = boto3.client("s3")

    def upload_file(self, local_path, url):
        _, b, k = split_url(url)
        self._s3.upload_file(local_path, b, k)

    def download_file(self, url, local_path):
        _, b, k = split_url(url)
        self._s3.download_file(b, k, local_path)

    def list(self, url_prefix):
        sch, b, k = split_url(url_prefix)
        out, token = [], None
        while True:
            kw = dict(Bucket=b, Prefix=k.rstrip("/") + "/")
            if token:
                kw["ContinuationToken"] = token
            r = self._s3.list_objects_v2(**kw)
            out += [f"{sch}://{b}/{o['Key']}" for o in r.get("Contents", [])]
            if not r.get("IsTruncated"):
                return out
            token = r.get("NextContinuationToken")

    def exists(self, url):
        _, b, k = split_url(url)
        r = self._s3.list_objects_v2(Bucket=b, Prefix=k, MaxKeys=1)
        return any(o["Key"] == k for o in r.get("Contents", []))

    def remove_prefix(self, url_prefix):
        _, b, _ = split_url(url_prefix)
        keys = [split_url(u)[2] for u in self.list(url_prefix)]
        for i in range(0, len(keys), 1000):
            self._s3.delete_objects(Bucket=b, Delete={"Objects": [{"Key": k} for k in keys[i:i + 1000]]})


class DirectoryStore(ObjectStore):
    """Object store emulated on a directory (``scheme://bucket/key`` -> ``<root>/bucket/key``).  Used for tests and for
    shared-filesystem "buckets"."""

    def __init__(self, root: str):
        self.root = root

    def _p(self, url):
        _, b, k = split_url(url)
        return os.path.join(self.root, b, *[s for s in k.split("/") if s])

    def upload_file(self, local_path, url):
        p = self._p(url)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        tmp = p + ".part"
        shutil.copyfile(local_path, tmp)
        os.replace(tmp, p)

    def download_file(self, url, local_path):
        shutil.copyfile(self._p(url), local_path)

    def list(self, url_prefix):
        sch, b, k = split_url(url_prefix)
        base = self._p(url_prefix)
        out = []
        for dirpath, _, files in os.walk(base):
            for f in files:
                if f.endswith(".part"):
                    continue
                rel = os.path.relpath(os.path.join(dirpath, f), os.path.join(self.root, b)).replace(os.sep, "/")
                out.append(f"{sch}://{b}/{rel}")
        return sorted(out)

    def exists(self, url):
        return os.path.isfile(self._p(url))

    def remove_prefix(self, url_prefix):
        shutil.rmtree(self._p(url_prefix), ignore_errors=True)


_FACTORIES: Dict[str, Callable[[], ObjectStore]] = {"s3": S3Store}
_STORES: Dict[str, ObjectStore] = {}


def register_object_store(scheme: str, factory: Callable[[], ObjectStore]) -> None:
    _FACTORIES[scheme] = factory
    _STORES.pop(scheme, None)


def get_store(url: str) -> ObjectStore:
    scheme = url.split("://", 1)[0]
    if scheme not in _STORES:
        if scheme not in _FACTORIES:
            raise ValueError(f"no object-store back end registered for '{scheme}://' paths")
        _STORES[scheme] = _FACTORIES[scheme]()
    return _STORES[scheme]


def staging_dir(url: str) -> str:
    """Node-local staging directory for a remote checkpoint root (``NXDT_CKPT_STAGING`` overrides the base)."""
    base = os.environ.get("NXDT_CKPT_STAGING", os.path.join(os.environ.get("TMPDIR", "/tmp"), "nxdt_ckpt_staging"))
    d = os.path.join(base, hashlib.sha1(url.encode()).hexdigest()[:16])
    os.makedirs(d, exist_ok=True)
    return d
