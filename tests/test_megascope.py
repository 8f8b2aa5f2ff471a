"""MegaScope: flags, whitelisted compressor, capture through a real model, perturbations,
WebSocket protocol against a live server."""
import json
import threading
import time

import pytest
import torch

from megatronapp_b200.scope import hooks
from megatronapp_b200.scope.capture import DefaultCompressor, FlagType, TensorTracers, TTFlags, resolve_method
from megatronapp_b200.scope.disturbance import Disturbance


def test_flag_codes_and_defaults():
    assert [f.value for f in FlagType] == list(range(8))
    f = TTFlags(3)
    f.set_by_configs({"QKV_mat_mul": "true"})
    assert f.get_flag(FlagType.QKV_mat_mul, 2) and f.get_flag(FlagType.MLP1_mat_mul, 1)   # MLP1 defaults to True
    assert not f.get_flag(FlagType.RawAttentionScore_mat_mul, 1)


def test_compressor_is_a_whitelist_not_eval():
    assert resolve_method("data.mean(dim=-1)") == "mean"
    assert resolve_method("data.abs().max(dim=-1).values") == "absmax"
    with pytest.raises(ValueError):
        resolve_method("__import__('os').system('id')")
    c = DefaultCompressor()
    with pytest.raises(ValueError):
        c.set_by_configs({"QKV": {"pixels": 8, "method": "open('/etc/passwd').read()"}})
    x = torch.arange(2 * 3 * 16, dtype=torch.float32).view(2, 3, 16)
    out = c.compress_tensor(x, 4, "mean")
    assert out.shape == (2, 3, 4) and torch.allclose(out[0, 0], torch.tensor([1.5, 5.5, 9.5, 13.5]))


def test_capture_through_model_and_perturbation():
    from megatronapp_b200.bench_model import build_gpt_for_bench, synthetic_batch, train_step

    st = build_gpt_for_bench(2, 64, 4, 32, 256, 2, device="cpu", dtype=torch.float32)
    b = synthetic_batch(st, "cpu")
    tr = TensorTracers(2)
    got = []
    tr.set_report(lambda name, args, data: got.append((name[0], int(name[1]), args, len(data))))
    tr.flags.set_by_configs({"QKV_mat_mul": "true", "RawAttentionScore_mat_mul": "true",
                             "ContextLayer_mat_mul": "true", "MLP2_Plot": "true"})
    tr.compressor.set_by_configs({"QKV": {"pixels": 12, "method": "mean"}, "MLP": {"pixels": 8, "method": "absmax"}})
    dis = Disturbance()
    hooks.install(tr, dis)
    try:
        base = train_step(st, [b], do_optimizer=False)
        kinds = {(l, k) for l, k, _, _ in got}
        assert {(1, 1), (1, 2), (1, 3), (1, 4), (1, 5), (2, 1), (2, 5)} <= kinds
        qkv = [g for g in got if g[1] == 1][0]
        assert qkv[2] == [32] and qkv[3] == 2 * 32 * 12            # [b, s, px] flattened
        attn = [g for g in got if g[1] == 2][0]
        assert attn[2] == [4, 32, 32]
        tr.tik_end()
        assert any(k == 7 for _, k, _, _ in got)                     # PCA plot records
        tr.flags.clear()
        dis.set_by_configs({"system_perturbation": True, "system_perturbation_fn": "noise2",
                            "system_perturbation_coef": 0.5})
        noisy = train_step(st, [b], do_optimizer=False)
        assert abs(noisy - base) > 1e-4
    finally:
        hooks.install(None, None)
    assert abs(train_step(st, [b], do_optimizer=False) - base) < 1e-6   # hooks off => identical


def test_training_ws_protocol():
    from websockets.sync.client import connect

    from megatronapp_b200.scope.training_server import TrainingWSServer

    srv = TrainingWSServer(port=18765, num_layers=2, micro_batch_size=2, seq_length=32)
    srv.start()
    try:
        time.sleep(0.3)
        with connect("ws://127.0.0.1:18765") as ws:
            ws.send(json.dumps({"type": "ping"}))
            assert json.loads(ws.recv(timeout=5)) == {"type": "pong"}
            ws.send(json.dumps({"type": "run_training_step", "visualization_flags": {"QKV_mat_mul": "true"},
                                "compressor_config": {"QKV": {"pixels": 4, "method": "data.mean(dim=-1)"}}}))
            t = threading.Thread(target=srv.wait_for_step_request)
            t.start(); t.join(5)
            assert json.loads(ws.recv(timeout=5)) == {"type": "start", "micro_batch_size": 2, "seq_length": 32, "num_layers": 2}
            assert srv.tracers.flags.get_flag(FlagType.QKV_mat_mul, 1)
            srv.tracers.tik_tensor((1, 1), torch.randn(32, 2, 24))
            msg = json.loads(ws.recv(timeout=5))
            assert msg["type"] == "update" and msg["update_type"] == 1 and msg["layer_id"] == 1 and msg["args"] == [32]
            assert len(msg["result"]) == 2 * 32 * 4
    finally:
        srv.shutdown()


def test_frontend_modules_resolve_and_are_balanced():
    """The component SPA ships without a build step: every ES-module import must resolve to a file in the tree, the
    entry point must mount every view of the reference UI, and brackets must balance (cheap syntax guard: no JS runtime
    in the CI image)."""
    import os
    import re

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "megatronapp_b200", "scope", "ui")
    html = open(os.path.join(root, "index.html")).read()
    assert 'type="module" src="js/main.js"' in html and 'href="style.css"' in html
    seen, todo = set(), [os.path.join(root, "js", "main.js")]
    while todo:
        path = todo.pop()
        if path in seen:
            continue
        seen.add(path)
        src = open(path).read()
        code = re.sub(r"`(?:[^`\\]|\\.)*`|\"(?:[^\"\\\n]|\\.)*\"|'(?:[^'\\\n]|\\.)*'", '""', src)   # strings first ("ws://...")
        code = re.sub(r"//[^\n]*", "", code)
        for a, b in ("()", "[]", "{}"):
            assert code.count(a) == code.count(b), (path, a, code.count(a), code.count(b))
        for m in re.finditer(r'from\s+"(\.[^"]+)"', src):
            dep = os.path.normpath(os.path.join(os.path.dirname(path), m.group(1)))
            assert os.path.isfile(dep), (path, m.group(1))
            todo.append(dep)
    names = {os.path.basename(p) for p in seen}
    assert {"ControlPanel.js", "TokenStrip.js", "Selectors.js", "VectorViews.js", "ColoredVector.js", "AttentionMatrix.js",
            "OutputProbs.js", "PCAPlot.js", "store.js", "protocol.js", "draw.js"} <= names, names
    proto = open(os.path.join(root, "js", "protocol.js")).read()
    for word in ("generate", "run_training_step", "ping", "start", "update", "finish", "error"):
        assert word in proto
