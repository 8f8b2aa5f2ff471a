#!/usr/bin/env python
"""Tiny undefined-name check (a pyflakes subset, no dependencies): every ``Name`` that is read must be bound somewhere in
an enclosing function, the module, or builtins.  Catches NameErrors in branches the CPU test-suite cannot execute
(CUDA-only paths).  ``python tools/check_names.py [paths...]`` -> exit 1 on findings."""
import ast
import builtins
import os
import sys


def _bound_names(node) -> set:
    out = set()
    for n in ast.walk(node):
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            out.add(n.name)
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
            a = n.args
            for x in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
                out.add(x.arg)
        elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            out.add(n.id)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            for al in n.names:
                out.add((al.asname or al.name).split(".")[0])
        elif isinstance(n, ast.ExceptHandler) and n.name:
            out.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            out.update(n.names)
        elif isinstance(n, ast.MatchAs) and n.name:
            out.add(n.name)
    return out


def check(path: str):
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src, path)
    if any(isinstance(n, ast.ImportFrom) and any(a.name == "*" for a in n.names) for n in ast.walk(tree)):
        return []
    known = _bound_names(tree) | set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__class__", "__spec__", "__package__"}
    return [f"{path}:{n.lineno}: undefined name {n.id!r}" for n in ast.walk(tree)
            if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in known]


def main(argv):
    roots = argv or ["megatronapp_b200", "tools", "tasks", "bench.py", "__graft_entry__.py"] + \
        [f for f in os.listdir(".") if f.startswith("pretrain_") and f.endswith(".py")]
    findings = []
    for r in roots:
        files = [r] if os.path.isfile(r) else [os.path.join(d, f) for d, _, fs in os.walk(r) for f in fs if f.endswith(".py")]
        for f in files:
            if "_ref" not in f:
                findings += check(f)
    print("\n".join(findings) if findings else "names clean")
    return 1 if findings else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
