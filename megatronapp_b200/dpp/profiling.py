"""Post-processing of MegaScan traces for pipeline-schedule studies (static interleaved 1F1B vs
MegaDPP): iteration time, "reduce window" (time from the ``reduce`` marker -- the moment the last
gradient of a chunk is produced -- to the end of the iteration), computing / sending ratios and
windows per pipeline rank, and peak memory.

Parity: reference ``profiling/process_iter_times.py:15-43`` (iteration time + reduce window),
``process_computing_{ratio,window}*.py``, ``process_sending_{ratio,window}*.py``,
``process_peak_memory.py`` and the root ``aggregate.py`` trace merger.  Input is either a
directory of per-rank ``benchmark-data-*.json`` files or an already aggregated Chrome trace.
"""
from __future__ import annotations

import json
import math
import os
from collections import defaultdict
from typing import Dict, List, Optional

from ..scan.aggregate import benchmark_to_chrome_trace, read_benchmark_dir

COMPUTE_NAMES = ("forward", "backward")
SEND_NAMES = ("send_forward", "send_backward", "send_next", "send_prev", "p2p_send", "mailbox_put")
RECV_NAMES = ("recv_forward", "recv_backward", "recv_next", "recv_prev", "p2p_recv", "mailbox_get")


def load_trace(path: str) -> List[dict]:
    if os.path.isdir(path):
        return benchmark_to_chrome_trace(read_benchmark_dir(path))
    with open(path) as f:
        data = json.load(f)
    events = data["traceEvents"] if isinstance(data, dict) else data
    return normalize_marker_trace(events)


def normalize_marker_trace(events: List[dict]) -> List[dict]:
    """The reference's DPP tracer (``core/trace.py``) writes B/E pairs plus ``iteration begin`` /
    ``iteration end`` instant markers instead of complete events; convert that dialect (its committed
    ``profiling/*-aggregated.json`` files) into ``X`` events with one ``iteration`` span per rank."""
    if not any(e.get("name") == "iteration begin" for e in events):
        return events
    out: List[dict] = []
    open_ev = {}
    begins = {}
    for e in sorted(events, key=lambda x: x["ts"]):
        pid, name, ph = e.get("pid", 0), e.get("name"), e.get("ph")
        if name == "iteration begin":
            begins[pid] = e
        elif name == "iteration end":
            b = begins.pop(pid, None)
            if b is not None:
                out.append({"name": "iteration", "ph": "X", "ts": b["ts"], "dur": e["ts"] - b["ts"], "pid": pid,
                            "tid": 0, "args": dict(b.get("args", {}))})
        elif ph == "B":
            open_ev.setdefault((pid, name), []).append(e)
        elif ph == "E":
            st = open_ev.get((pid, name))
            if st:
                b = st.pop()
                out.append({"name": name, "ph": "X", "ts": b["ts"], "dur": e["ts"] - b["ts"], "pid": pid, "tid": 0,
                            "args": {**b.get("args", {}), **e.get("args", {})}})
        else:
            out.append(dict(e))
    out.sort(key=lambda x: x["ts"])
    return out


def _union_length(intervals):
    tot, end = 0.0, -math.inf
    for b, e in sorted(intervals):
        if e <= end:
            continue
        tot += e - max(b, end)
        end = e
    return tot


def _iterations(events: List[dict]) -> Dict[int, List[dict]]:
    """pid -> list of {'begin','end','events'} using the per-iteration top-level span."""
    per = defaultdict(list)
    for e in events:
        if e.get("ph") in ("X", "i"):
            per[e["pid"]].append(e)
    out = {}
    for pid, evs in per.items():
        evs.sort(key=lambda e: e["ts"])
        spans = [e for e in evs if e["name"] == "iteration" and e["ph"] == "X"]
        its = []
        if spans:
            for k, sp in enumerate(spans):
                b, en = sp["ts"], sp["ts"] + sp.get("dur", 0.0)
                g = [e for e in evs if e is not sp and b - 1e-6 <= e["ts"] <= en + 1e-6]
                its.append({"begin": b, "end": en, "events": g, "iteration": sp.get("args", {}).get("iteration", k)})
        elif evs:
            its.append({"begin": evs[0]["ts"], "end": max(x["ts"] + x.get("dur", 0.0) for x in evs), "events": evs,
                        "iteration": 0})
        out[pid] = its
    return out


def analyze(events: List[dict], skip_first: int = 1) -> Dict[str, object]:
    """Per-rank means over iterations (the first ``skip_first`` are warm-up)."""
    res = {"ranks": {}}
    for pid, its in sorted(_iterations(events).items()):
        its = its[skip_first:] if len(its) > skip_first else its
        it_ms, red_win, comp_ratio, comp_win, send_ratio, send_win, peak = [], [], [], [], [], [], []
        for it in its:
            dur = it["end"] - it["begin"]
            if dur <= 0:
                continue
            it_ms.append(dur / 1e3)
            comp = [(e["ts"], e["ts"] + e["dur"]) for e in it["events"] if e["ph"] == "X" and e["name"] in COMPUTE_NAMES]
            send = [(e["ts"], e["ts"] + e["dur"]) for e in it["events"] if e["ph"] == "X" and e["name"] in SEND_NAMES]
            reduces = [e["ts"] for e in it["events"] if e["name"] == "reduce"]
            if reduces:
                red_win.append((it["end"] - min(reduces)) / 1e3)
            if comp:
                comp_ratio.append(_union_length(comp) / dur)
                comp_win.append((max(e for _, e in comp) - min(b for b, _ in comp)) / 1e3)
            if send:
                send_ratio.append(_union_length(send) / dur)
                send_win.append((max(e for _, e in send) - min(b for b, _ in send)) / 1e3)
            mems = [e["args"].get("peak_mem") for e in it["events"] if e.get("args", {}).get("peak_mem") is not None]
            if mems:
                peak.append(max(mems))
        mean = lambda xs: sum(xs) / len(xs) if xs else None
        gmean = lambda xs: math.exp(sum(math.log(max(x, 1e-12)) for x in xs) / len(xs)) if xs else None
        res["ranks"][pid] = {"iteration_ms": mean(it_ms), "reduce_window_ms": mean(red_win),
                             "computing_ratio": gmean(comp_ratio), "computing_window_ms": mean(comp_win),
                             "sending_ratio": gmean(send_ratio), "sending_window_ms": mean(send_win),
                             "peak_memory_bytes": max(peak) if peak else None, "iterations": len(it_ms)}
    vals = [r["iteration_ms"] for r in res["ranks"].values() if r["iteration_ms"]]
    res["iteration_ms"] = max(vals) if vals else None
    return res


def compare(static_path: str, dpp_path: str, skip_first: int = 1) -> Dict[str, object]:
    a, b = analyze(load_trace(static_path), skip_first), analyze(load_trace(dpp_path), skip_first)
    out = {"static": a, "dpp": b}
    if a["iteration_ms"] and b["iteration_ms"]:
        out["dpp_speedup"] = a["iteration_ms"] / b["iteration_ms"]
    return out


def plot(cmp: Dict[str, object], out_dir: str) -> List[str]:
    """Grouped bar charts per metric (needs matplotlib; silently skipped when absent)."""
    try:
        import matplotlib

        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:
        return []
    os.makedirs(out_dir, exist_ok=True)
    files = []
    ranks = sorted(cmp["static"]["ranks"])
    for metric in ("reduce_window_ms", "computing_ratio", "computing_window_ms", "sending_ratio", "sending_window_ms"):
        s = [cmp["static"]["ranks"][r][metric] or 0 for r in ranks]
        d = [cmp["dpp"]["ranks"].get(r, {}).get(metric) or 0 for r in ranks]
        x = range(len(ranks))
        plt.figure(figsize=(8, 4.5))
        plt.bar([i - 0.18 for i in x], s, 0.36, label="static 1F1B")
        plt.bar([i + 0.18 for i in x], d, 0.36, label="MegaDPP")
        plt.xticks(list(x), [str(r) for r in ranks])
        plt.xlabel("global rank")
        plt.ylabel(metric)
        plt.legend()
        plt.tight_layout()
        f = os.path.join(out_dir, f"{metric}.png")
        plt.savefig(f)
        plt.close()
        files.append(f)
    return files


def main(argv=None):
    import argparse

    ap = argparse.ArgumentParser(description="MegaDPP trace analysis")
    ap.add_argument("--trace", help="trace dir / aggregated json (single run)")
    ap.add_argument("--static", help="trace of the static schedule")
    ap.add_argument("--dpp", help="trace of the dynamic schedule")
    ap.add_argument("--plots", default=None, help="directory for bar charts")
    ap.add_argument("--skip-first", type=int, default=1)
    a = ap.parse_args(argv)
    if a.trace:
        out = analyze(load_trace(a.trace), a.skip_first)
    else:
        out = compare(a.static, a.dpp, a.skip_first)
        if a.plots:
            out["plots"] = plot(out, a.plots)
    print(json.dumps(out, indent=1, default=str))
    return out


if __name__ == "__main__":
    main()
