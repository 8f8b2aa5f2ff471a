#!/usr/bin/env python
"""Repository lint (reference ``tools/linter.py`` role): syntax-check every Python file, flag tabs / trailing
whitespace / over-long lines in Python, CUDA and C++ sources.  Exit code 1 on any finding.  ``python tools/linter.py [paths...]``"""
import ast
import os
import sys

MAX_LEN = 160


def lint_file(path: str):
    out = []
    try:
        src = open(path, encoding="utf-8").read()
    except UnicodeDecodeError:
        return [f"{path}: not utf-8"]
    if path.endswith(".py"):
        try:
            ast.parse(src, path)
        except SyntaxError as e:
            out.append(f"{path}:{e.lineno}: syntax error: {e.msg}")
    for i, line in enumerate(src.split("\n"), 1):
        if "\t" in line and path.endswith(".py"):
            out.append(f"{path}:{i}: tab character")
        if line != line.rstrip():
            out.append(f"{path}:{i}: trailing whitespace")
        if len(line) > MAX_LEN and "http" not in line:
            out.append(f"{path}:{i}: line longer than {MAX_LEN}")
    return out


def main(argv):
    roots = argv or ["megatronapp_b200", "tools", "tasks", "tests", "bench.py", "__graft_entry__.py"]
    findings = []
    for r in roots:
        if os.path.isfile(r):
            findings += lint_file(r)
            continue
        for d, _, files in os.walk(r):
            if "_ref" in d or "__pycache__" in d:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                    findings += lint_file(os.path.join(d, f))
    print("\n".join(findings) if findings else "lint clean")
    return 1 if findings else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
