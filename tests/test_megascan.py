"""MegaScan: tracer record schema, aggregator -> Chrome trace, dependency linking, detector."""
import json
import os

import torch

from megatronapp_b200.scan import aggregate as agg
from megatronapp_b200.scan.tracer import Tracer


def _fake_rank(dp, pp, tp, g_rk, slow=1.0, iters=3):
    recs = []
    for it in range(iters):
        t = 0
        recs.append({"name": "iteration", "ph": "B", "rel_ts": 0, "iteration": it, "pad_before": 0,
                     "unix_ns": 10 ** 9 * it, "dp_rk": dp, "pp_rk": pp, "tp_rk": tp, "g_rk": g_rk, "dev": g_rk})
        info = {"dp_rk": dp, "pp_rk": pp, "tp_rk": tp, "g_rk": g_rk, "dev": g_rk}
        for mb in range(4):
            for name, base in (("forward", 1000), ("backward", 2000)):
                dur = int(base * 1000 * slow)
                recs.append({"name": name, "ph": "B", "rel_ts": t, **info})
                if name == "forward":
                    # an all-reduce inside: the slow rank arrives last => waits least
                    ar = 50_000 if slow > 1.0 else int(50_000 + 1000 * 1000 * 0.3)
                    recs.append({"name": "_reduce", "ph": "B", "rel_ts": t + 10, "group": [0, 1, 2, 3],
                                 "expect": "_reduce", **info})
                    recs.append({"name": "_reduce", "ph": "E", "rel_ts": t + 10 + ar, **info})
                t += dur
                recs.append({"name": name, "ph": "E", "rel_ts": t, **info})
        wait = 10_000 if slow > 1.0 else 900_000
        recs.append({"name": "allreduce", "ph": "B", "rel_ts": t, **info})
        t += wait
        recs.append({"name": "allreduce", "ph": "E", "rel_ts": t, **info})
        recs.append({"name": "iteration", "ph": "E", "rel_ts": t, "iteration": it, "duration_wall": t,
                     "duration_cuda": t, **info})
    return recs


def test_aggregate_and_detect(tmp_path):
    d = tmp_path / "bench"
    d.mkdir()
    for dp in range(4):
        recs = _fake_rank(dp, 0, 0, dp, slow=1.35 if dp == 2 else 1.0)
        (d / f"benchmark-data-{dp}-pipeline-0-tensor-0.json").write_text(json.dumps(recs))
    out = tmp_path / "agg.json"
    ab = tmp_path / "abnormal.txt"
    traces, findings = agg.aggregate(str(d), str(out), detect=True, abnormal_file=str(ab))
    loaded = json.loads(out.read_text())
    assert any(e["ph"] == "X" and e["name"] == "forward" for e in loaded)
    assert {e["args"]["name"] for e in loaded if e.get("ph") == "M" and e["name"] == "process_name"} == {
        f"DP{i}-PP0-TP0" for i in range(4)}
    red = [e for e in loaded if e["name"] == "_reduce"]
    assert all(len(e["args"]["related_sync_op"].split()) == 4 for e in red)
    assert ab.read_text().strip() == "2 2-0-0"
    assert findings[0]["g_rk"] == 2 and findings[0]["confirmed"]


def test_healthy_run_reports_nothing(tmp_path):
    d = tmp_path / "bench"
    d.mkdir()
    for dp in range(4):
        (d / f"benchmark-data-{dp}-pipeline-0-tensor-0.json").write_text(json.dumps(_fake_rank(dp, 0, 0, dp)))
    ab = tmp_path / "abnormal.txt"
    _, findings = agg.aggregate(str(d), str(tmp_path / "o.json"), detect=True, abnormal_file=str(ab))
    assert not ab.exists() and not [f for f in findings if f["confirmed"]]


def _lockstep_tp_rank(tp, slow_rank, n_tp=4, iters=3, layers=6):
    """A TP group in lock step: every layer = compute + one collective; the slow rank computes 30 % longer and the others
    spend exactly that extra time waiting inside the collective -> identical ``forward`` durations on every rank."""
    info = {"dp_rk": 0, "pp_rk": 0, "tp_rk": tp, "g_rk": tp, "dev": tp}
    recs = []
    for it in range(iters):
        t = 0
        recs.append({"name": "iteration", "ph": "B", "rel_ts": 0, "iteration": it, "pad_before": 0, **info})
        for name in ("forward", "backward"):
            recs.append({"name": name, "ph": "B", "rel_ts": t, **info})
            for _ in range(layers):
                comp = 1_300_000 if tp == slow_rank else 1_000_000
                t += comp
                recs.append({"name": "gemm_rs", "ph": "B", "rel_ts": t, "group": list(range(n_tp)), "expect": "gemm_rs", **info})
                t += 200_000 + (1_300_000 - comp)              # transfer + waiting for the slow peer
                recs.append({"name": "gemm_rs", "ph": "E", "rel_ts": t, **info})
            recs.append({"name": name, "ph": "E", "rel_ts": t, **info})
        recs.append({"name": "iteration", "ph": "E", "rel_ts": t, "iteration": it, "duration_wall": t, "duration_cuda": t, **info})
    return recs


def test_detects_slow_rank_inside_a_lockstep_tensor_parallel_group(tmp_path):
    d = tmp_path / "bench"
    d.mkdir()
    for tp in range(4):
        (d / f"benchmark-data-0-pipeline-0-tensor-{tp}.json").write_text(json.dumps(_lockstep_tp_rank(tp, slow_rank=1)))
    ab = tmp_path / "abnormal.txt"
    _, findings = agg.aggregate(str(d), str(tmp_path / "o.json"), detect=True, abnormal_file=str(ab))
    assert ab.read_text().strip() == "1 0-0-1"
    f = findings[0]
    assert f["g_rk"] == 1 and f["confirmed"] and f["sync_all"] > 0 and f["sync_time_ratio"] < 0.5
    # and a healthy lock-step group reports nothing
    for tp in range(4):
        (d / f"benchmark-data-0-pipeline-0-tensor-{tp}.json").write_text(json.dumps(_lockstep_tp_rank(tp, slow_rank=-1)))
    ab.unlink()
    _, findings = agg.aggregate(str(d), str(tmp_path / "o.json"), detect=True, abnormal_file=str(ab))
    assert not ab.exists() and not [x for x in findings if x["confirmed"]]


def test_amend_p2p():
    tr = [{"name": "send-forward", "ph": "X", "ts": 0.0, "dur": 100.0, "pid": 0, "args": {"peer": 1, "data": 1000, "id": 0}},
          {"name": "recv-forward", "ph": "X", "ts": 60.0, "dur": 45.0, "pid": 1, "args": {"peer": 0, "data": 1000, "id": 1}}]
    agg.amend_p2p(tr)
    assert tr[0]["dur"] == tr[1]["dur"] == 45.0 and tr[0]["ts"] == tr[1]["ts"] == 55.0


def test_tracer_cpu_roundtrip(tmp_path):
    t = Tracer()
    t.configure(True, str(tmp_path), interval=2, continuous_iterations=1)
    for it in range(4):
        t.iteration_begin(it)
        with t.scope("forward", microbatch=0):
            with t.scope("_reduce", group=lambda: [0], expect="_reduce", data=4096):
                torch.zeros(10).sum()
        t.iteration_end()
        t.iteration_end()  # idempotent (reference called it twice and duplicated records)
    t.shutdown()
    f = tmp_path / "benchmark-data-0-pipeline-0-tensor-0.json"
    recs = json.loads(f.read_text())
    its = [r for r in recs if r["name"] == "iteration" and r["ph"] == "B"]
    assert [r["iteration"] for r in its] == [0, 2]          # traced 1 of every 2 iterations
    assert sum(1 for r in recs if r["name"] == "forward" and r["ph"] == "B") == 2
    red = [r for r in recs if r["name"] == "_reduce" and r["ph"] == "B"][0]
    assert red["group"] == [0] and red["data"] == 4096 and "bandwidth" in red
    traces = agg.benchmark_to_chrome_trace({(0, 0, 0): recs})
    assert any(e["name"] == "forward" and e["ph"] == "X" for e in traces)


def test_bench_scope_digest_counts_outermost_comm_only():
    """bench.py's exposed-comm metric: nested comm scopes are counted once, compute scopes never."""
    import bench

    recs = [
        {"name": "iteration", "ph": "B", "rel_ts": 0},
        {"name": "forward", "ph": "B", "rel_ts": 10},
        {"name": "_reduce_scatter_along_first_dim", "ph": "B", "rel_ts": 20},
        {"name": "_reduce_scatter_along_first_dim", "ph": "E", "rel_ts": 50},
        {"name": "forward", "ph": "E", "rel_ts": 100},
        {"name": "allreduce", "ph": "B", "rel_ts": 110},
        {"name": "grad-reduce-scatter", "ph": "B", "rel_ts": 120},
        {"name": "grad-reduce-scatter", "ph": "E", "rel_ts": 150},
        {"name": "allreduce", "ph": "E", "rel_ts": 160},
        {"name": "iteration", "ph": "E", "rel_ts": 200},
    ]
    d = bench._digest_scopes(recs)
    assert d["exposed_comm_ns"] == 30 + 50
    assert d["per_scope"]["forward"] == {"ns": 90, "calls": 1}
    assert d["per_scope"]["grad-reduce-scatter"]["ns"] == 30
