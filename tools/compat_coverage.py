"""Coverage of the reference's Python API by the ``compat/megatron`` namespace.

For every module under ``<reference>/megatron`` reports whether ``import megatron.<...>`` resolves here and which of its public
top-level classes / functions are missing from the module it resolves to.  ``python tools/compat_coverage.py /root/reference``"""
import ast
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def public_names(path):
    try:
        tree = ast.parse(open(path, encoding="utf-8").read())
    except SyntaxError:
        return []
    return [n.name for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and not n.name.startswith("_")]


def main(ref):
    sys.path[:0] = [os.path.join(ROOT, "compat"), ROOT]
    import megatron  # noqa: F401

    rows, tot_names, have_names, resolved = [], 0, 0, 0
    for d, _, files in os.walk(os.path.join(ref, "megatron")):
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            path = os.path.join(d, f)
            mod = os.path.relpath(path, ref)[:-3].replace(os.sep, ".")
            if mod.endswith(".__init__"):
                mod = mod[:-9]
            if ".tests" in mod or mod.endswith(".setup"):
                continue
            names = public_names(path)
            try:
                m = importlib.import_module(mod)
            except Exception:  # noqa: BLE001
                rows.append({"module": mod, "resolves": False, "missing": names})
                tot_names += len(names)
                continue
            resolved += 1
            missing = [n for n in names if not hasattr(m, n)]
            tot_names += len(names)
            have_names += len(names) - len(missing)
            rows.append({"module": mod, "resolves": True, "missing": missing})
    summary = {"modules": len(rows), "modules_resolved": resolved, "public_names": tot_names, "public_names_available": have_names}
    return summary, rows


if __name__ == "__main__":
    summary, rows = main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print(json.dumps(summary))
    if "-v" in sys.argv:
        for r in rows:
            if not r["resolves"] or r["missing"]:
                print(("MISSING MODULE " if not r["resolves"] else "partial        ") + r["module"], r["missing"][:12])
