"""YAML config system: attribute-access dict tree, ``${a.b}`` / ``${.rel}`` interpolation,
``${multiply:a,b}`` style resolvers and ``key=value`` CLI overrides.

The reference relies on hydra + omegaconf for this (reference:
examples/training_orchestrator.py:140-145, examples/conf/hf_llama3_8B_config.yaml:33 for the
``multiply`` resolver).  Neither library is available offline, and we only need a small, fast
subset, so this is a self-contained implementation with lazy interpolation (values are resolved on
access, like omegaconf, so CLI overrides applied after load are visible through interpolations).
"""
from __future__ import annotations

import copy
import os
import re
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional

import yaml

_RESOLVERS: Dict[str, Callable[..., Any]] = {}


def register_resolver(name: str, fn: Callable[..., Any]) -> None:
    _RESOLVERS[name] = fn


def _num(x):
    if isinstance(x, (int, float)):
        return x
    s = str(x).strip()
    try:
        return int(s)
    except ValueError:
        return float(s)


register_resolver("multiply", lambda a, b: _num(a) * _num(b))
register_resolver("add", lambda a, b: _num(a) + _num(b))
register_resolver("divide", lambda a, b: _num(a) / _num(b))
register_resolver("floor_div", lambda a, b: _num(a) // _num(b))
register_resolver("oc.env", lambda k, d=None: os.environ.get(str(k), d))
register_resolver("env", lambda k, d=None: os.environ.get(str(k), d))

_MISSING = object()


class ConfigError(KeyError):
    pass


class Config:
    """A node of the config tree (mapping).  Lists are kept as ``ConfigList``."""

    __slots__ = ("_data", "_parent", "_key")

    def __init__(self, data: Optional[dict] = None, parent: "Optional[Config | ConfigList]" = None, key=None):
        object.__setattr__(self, "_data", {})
        object.__setattr__(self, "_parent", parent)
        object.__setattr__(self, "_key", key)
        if data:
            for k, v in data.items():
                self._data[k] = _wrap(v, self, k)

    # ---- tree helpers -------------------------------------------------------------------------
    def _root(self):
        node = self
        while node._parent is not None:
            node = node._parent
        return node

    def _resolve(self, value):
        if isinstance(value, str) and "${" in value:
            return _interpolate(value, self)
        return value

    # ---- mapping protocol ---------------------------------------------------------------------
    def __getattr__(self, name: str):
        if name.startswith("__"):
            raise AttributeError(name)
        try:
            return self[name]
        except KeyError:
            raise AttributeError(f"config has no key '{name}'") from None

    def __setattr__(self, name: str, value):
        self[name] = value

    def __getitem__(self, key):
        if key not in self._data:
            raise ConfigError(key)
        return self._resolve(self._data[key])

    def __setitem__(self, key, value):
        self._data[key] = _wrap(value, self, key)

    def __delitem__(self, key):
        del self._data[key]

    def __contains__(self, key) -> bool:
        return key in self._data

    def __iter__(self) -> Iterator:
        return iter(self._data)

    def __len__(self) -> int:
        return len(self._data)

    def keys(self):
        return self._data.keys()

    def values(self):
        return [self[k] for k in self._data]

    def items(self):
        return [(k, self[k]) for k in self._data]

    def get(self, key, default=None):
        if key in self._data:
            v = self[key]
            return default if v is None and default is not None else v
        return default

    def pop(self, key, default=_MISSING):
        if key in self._data:
            v = self[key]
            del self._data[key]
            return v
        if default is _MISSING:
            raise ConfigError(key)
        return default

    def setdefault(self, key, default):
        if key not in self._data:
            self[key] = default
        return self[key]

    def update(self, other) -> None:
        for k, v in (other.items() if hasattr(other, "items") else other):
            self[k] = v

    def __repr__(self) -> str:
        return f"Config({self.to_dict(resolve=False)!r})"

    def __deepcopy__(self, memo):
        return Config(copy.deepcopy(self.to_dict(resolve=False), memo))

    def __getstate__(self):
        return self.to_dict(resolve=False)

    def __setstate__(self, state):
        object.__setattr__(self, "_data", {})
        object.__setattr__(self, "_parent", None)
        object.__setattr__(self, "_key", None)
        for k, v in state.items():
            self._data[k] = _wrap(v, self, k)

    # ---- conversions --------------------------------------------------------------------------
    def to_dict(self, resolve: bool = True) -> dict:
        out = {}
        for k, v in self._data.items():
            if isinstance(v, (Config, ConfigList)):
                out[k] = v.to_dict(resolve) if isinstance(v, Config) else v.to_list(resolve)
            else:
                out[k] = self._resolve(v) if resolve else v
        return out

    # ---- dotted access ------------------------------------------------------------------------
    def select(self, path: str, default=_MISSING):
        node: Any = self
        for part in path.split("."):
            if part == "":
                continue
            if isinstance(node, Config):
                if part not in node:
                    if default is _MISSING:
                        raise ConfigError(path)
                    return default
                node = node[part]
            elif isinstance(node, ConfigList):
                node = node[int(part)]
            else:
                if default is _MISSING:
                    raise ConfigError(path)
                return default
        return node

    def set_path(self, path: str, value, create: bool = True) -> None:
        parts = [p for p in path.split(".") if p]
        node = self
        for part in parts[:-1]:
            if part not in node or not isinstance(node._data[part], Config):
                if not create:
                    raise ConfigError(path)
                node[part] = {}
            node = node._data[part]
        node[parts[-1]] = value


class ConfigList:
    __slots__ = ("_data", "_parent", "_key")

    def __init__(self, data: Iterable, parent=None, key=None):
        self._parent = parent
        self._key = key
        self._data = []
        for i, v in enumerate(data):
            self._data.append(_wrap(v, self, i))

    def _root(self):
        node = self
        while node._parent is not None:
            node = node._parent
        return node

    def _resolve(self, value):
        if isinstance(value, str) and "${" in value:
            return _interpolate(value, self)
        return value

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._resolve(v) for v in self._data[i]]
        return self._resolve(self._data[i])

    def __setitem__(self, i, v):
        self._data[i] = _wrap(v, self, i)

    def __len__(self):
        return len(self._data)

    def __iter__(self):
        return (self._resolve(v) for v in self._data)

    def __contains__(self, item):
        return any(v == item for v in self)

    def __eq__(self, other):
        return list(self) == list(other)

    def append(self, v):
        self._data.append(_wrap(v, self, len(self._data)))

    def to_list(self, resolve: bool = True) -> list:
        out = []
        for v in self._data:
            if isinstance(v, Config):
                out.append(v.to_dict(resolve))
            elif isinstance(v, ConfigList):
                out.append(v.to_list(resolve))
            else:
                out.append(self._resolve(v) if resolve else v)
        return out

    def __repr__(self):
        return f"ConfigList({self.to_list(False)!r})"

    def __deepcopy__(self, memo):
        return ConfigList(copy.deepcopy(self.to_list(False), memo))


def _wrap(v, parent, key):
    if isinstance(v, Config):
        return Config(v.to_dict(resolve=False), parent, key)
    if isinstance(v, ConfigList):
        return ConfigList(v.to_list(resolve=False), parent, key)
    if isinstance(v, dict):
        return Config(v, parent, key)
    if isinstance(v, (list, tuple)):
        return ConfigList(v, parent, key)
    return v


# ---- interpolation --------------------------------------------------------------------------------

def _find_matching(s: str, start: int) -> int:
    """index of the '}' closing the '${' that starts at ``start``."""
    depth = 0
    i = start
    while i < len(s):
        if s.startswith("${", i):
            depth += 1
            i += 2
            continue
        if s[i] == "}":
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError(f"unbalanced interpolation in {s!r}")


def _split_args(s: str) -> List[str]:
    args, depth, cur = [], 0, []
    i = 0
    while i < len(s):
        if s.startswith("${", i):
            depth += 1
            cur.append("${")
            i += 2
            continue
        c = s[i]
        if c == "}":
            depth -= 1
        if c == "," and depth == 0:
            args.append("".join(cur).strip())
            cur = []
        else:
            cur.append(c)
        i += 1
    args.append("".join(cur).strip())
    return args


def _lookup(path: str, node) -> Any:
    if path.startswith("."):
        # relative: one dot = sibling of the current node's container
        dots = len(path) - len(path.lstrip("."))
        base = node
        for _ in range(dots - 1):
            if base._parent is None:
                raise ConfigError(path)
            base = base._parent
        rest = path[dots:]
    else:
        base = node._root()
        rest = path
    cur: Any = base
    for part in [p for p in rest.split(".") if p]:
        if isinstance(cur, Config):
            cur = cur[part]
        elif isinstance(cur, ConfigList):
            cur = cur[int(part)]
        else:
            raise ConfigError(path)
    return cur


def _eval_expr(expr: str, node) -> Any:
    expr = expr.strip()
    m = re.match(r"^([A-Za-z_][\w.]*):(.*)$", expr, flags=re.S)
    if m and m.group(1) in _RESOLVERS:
        args = [_interpolate(a, node) if "${" in a else _coerce_scalar(a) for a in _split_args(m.group(2))]
        return _RESOLVERS[m.group(1)](*args)
    if m and m.group(1) not in _RESOLVERS and ":" in expr and not expr.startswith("."):
        raise ConfigError(f"unknown resolver '{m.group(1)}'")
    return _lookup(expr, node)


def _coerce_scalar(s: str):
    try:
        return yaml.load(s, Loader=_Loader)
    except yaml.YAMLError:
        return s


def _interpolate(value: str, node) -> Any:
    # whole-string interpolation keeps the referenced type
    if value.startswith("${") and _find_matching(value, 0) == len(value) - 1:
        return _eval_expr(value[2:-1], node)
    out, i = [], 0
    while i < len(value):
        j = value.find("${", i)
        if j < 0:
            out.append(value[i:])
            break
        out.append(value[i:j])
        k = _find_matching(value, j)
        out.append(str(_eval_expr(value[j + 2:k], node)))
        i = k + 1
    return "".join(out)


# ---- public API -----------------------------------------------------------------------------------

class _Loader(yaml.SafeLoader):
    pass


# YAML 1.1 parses "1e-4" as a string; omegaconf treats it as a float.  Match that.
_Loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    re.compile(r"""^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                 |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                 |\.[0-9_]+(?:[eE][-+][0-9]+)?
                 |[-+]?\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$""", re.X),
    list("-+0123456789."),
)


def load_yaml(path: str) -> Config:
    with open(path, "r") as f:
        data = yaml.load(f, Loader=_Loader) or {}
    return Config(data)


def from_dict(d: dict) -> Config:
    return Config(copy.deepcopy(d))


def apply_overrides(cfg: Config, overrides: Iterable[str]) -> Config:
    """hydra-style ``a.b.c=value`` (also ``+a.b=value`` to add, ``~a.b`` to delete)."""
    for ov in overrides:
        ov = ov.strip()
        if not ov:
            continue
        if ov.startswith("~"):
            path = ov[1:]
            parent_path, _, leaf = path.rpartition(".")
            parent = cfg.select(parent_path) if parent_path else cfg
            if leaf in parent:
                del parent[leaf]
            continue
        if "=" not in ov:
            raise ValueError(f"bad override {ov!r}; expected key=value")
        key, val = ov.split("=", 1)
        key = key.lstrip("+")
        v = _coerce_scalar(val) if val != "" else None
        cfg.set_path(key, v)
    return cfg


def load_config(config_path: str, config_name: Optional[str] = None, overrides: Iterable[str] = ()) -> Config:
    """``load_config("conf", "hf_llama3_8B_config", ["trainer.devices=8"])`` or a direct file path."""
    if config_name is not None:
        name = config_name if config_name.endswith((".yaml", ".yml")) else config_name + ".yaml"
        path = os.path.join(config_path, name)
    else:
        path = config_path
    cfg = load_yaml(path)
    apply_overrides(cfg, overrides)
    return cfg


def get_attribute_from_cfg(cfg: Config, attr: str, default=None):
    """Exact dotted path first, then search-anywhere-by-leaf-name (depth-first), mirroring
    reference utils/utils.py:79-149.  Prefer explicit paths in new code."""
    try:
        return cfg.select(attr)
    except (KeyError, ValueError, IndexError):
        pass
    leaf = attr.split(".")[-1]

    def walk(node):
        if isinstance(node, Config):
            if leaf in node:
                return node[leaf]
            for k in node:
                sub = node._data[k]
                if isinstance(sub, (Config, ConfigList)):
                    r = walk(sub)
                    if r is not _MISSING:
                        return r
        elif isinstance(node, ConfigList):
            for sub in node._data:
                if isinstance(sub, (Config, ConfigList)):
                    r = walk(sub)
                    if r is not _MISSING:
                        return r
        return _MISSING

    r = walk(cfg)
    return default if r is _MISSING else r
