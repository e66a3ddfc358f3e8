#!/usr/bin/env python
"""Tiny static check (no third-party linter in the image): names that are read somewhere in a module but bound nowhere
in it (module scope, any enclosing function / class scope, comprehension targets, imports, builtins) -- i.e. NameErrors
waiting in rarely executed branches -- and imports that are never used.  Usage: python tools/lint_names.py [paths...]"""
import ast
import builtins
import sys
from pathlib import Path


def bound_names(tree):
    names = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__class__"}
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            names.add(node.name)
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
            a = node.args
            for arg in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
                names.add(arg.arg)
        elif isinstance(node, ast.Name) and isinstance(node.ctx, (ast.Store, ast.Del)):
            names.add(node.id)
        elif isinstance(node, (ast.Import, ast.ImportFrom)):
            for al in node.names:
                names.add((al.asname or al.name).split(".")[0])
        elif isinstance(node, ast.ExceptHandler) and node.name:
            names.add(node.name)
        elif isinstance(node, (ast.Global, ast.Nonlocal)):
            names.update(node.names)
        elif isinstance(node, ast.MatchAs) and node.name:
            names.add(node.name)
    return names


def check(path: Path):
    src = path.read_text()
    tree = ast.parse(src, str(path))
    star = any(isinstance(n, ast.ImportFrom) and any(a.name == "*" for a in n.names) for n in ast.walk(tree))
    bound = bound_names(tree)
    problems = []
    used = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.Name) and isinstance(node.ctx, ast.Load):
            used.add(node.id)
            if node.id not in bound and not star:
                problems.append((node.lineno, f"undefined name '{node.id}'"))
        elif isinstance(node, ast.Attribute):
            pass
    exported = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "__all__" for t in node.targets):
            try:
                exported = set(ast.literal_eval(node.value))
            except Exception:  # noqa: BLE001
                pass
    if path.name != "__init__.py":
        for node in tree.body:
            if isinstance(node, (ast.Import, ast.ImportFrom)):
                if isinstance(node, ast.ImportFrom) and node.module == "__future__":
                    continue
                for al in node.names:
                    n = (al.asname or al.name).split(".")[0]
                    if n != "*" and n not in used and n not in exported and "noqa" not in src.splitlines()[node.lineno - 1]:
                        problems.append((node.lineno, f"unused import '{n}'"))
    return problems


def main(argv):
    roots = [Path(a) for a in argv] or [Path("neuronx_distributed_training_b200"), Path("examples"), Path("tools"),
                                         Path("bench.py"), Path("__graft_entry__.py")]
    bad = 0
    for root in roots:
        files = [root] if root.is_file() else sorted(root.rglob("*.py"))
        for f in files:
            if "_ref" in f.parts:
                continue
            for line, msg in sorted(check(f)):
                print(f"{f}:{line}: {msg}")
                bad += 1
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
