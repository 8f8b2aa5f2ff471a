"""Static checks that also cover code the CPU suite cannot execute (CUDA-only branches)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_no_undefined_names(capsys):
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        assert _load("check_names").main([]) == 0, capsys.readouterr().out
    finally:
        os.chdir(cwd)


def test_every_python_file_parses_and_has_no_tabs():
    lint = _load("linter")
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "megatronapp_b200")):
        for f in files:
            if f.endswith(".py"):
                bad += [x for x in lint.lint_file(os.path.join(d, f)) if "syntax error" in x or "tab character" in x]
    assert not bad, bad
