"""MegaScan offline analysis: merge per-rank traces -> Chrome trace, link collective peers,
amend P2P pairs, detect the slow ("abnormal") GPU.

Capability parity with reference ``scripts/aggregate.py`` (read ``benchmark-data-*.json`` :56-129,
merge per iteration :142-239, Chrome trace with ``ph: "X"``, ``pid`` = global rank,
``process_name`` ``DP{d}-PP{p}-TP{t}`` :242-363, detector ``try_detect`` :399-489 writing
``abnormal.txt`` as ``"<global_rank> <dp>-<pp>-<tp>"`` :393-396) and ``scripts/dependency.py``
(``related_sync_op`` sets :26-51, ``amendP2P`` :54-86).  CLI: ``-b/--bench-dir -o/--output -d``.

Differences by design: the detector does not depend on events that are never emitted
(SURVEY Q5): it uses whatever of ``forward/backward/loss/allreduce`` + TP collectives exist in
the trace, works per (pp, tp) across DP *and* per (dp, pp) across TP, and optionally aligns
timelines with the wall-clock anchor recorded by the tracer (``unix_ns``).
"""
from __future__ import annotations

import argparse
import collections
import glob
import json
import logging
import os
import re
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

_FILE_RE = re.compile(r"benchmark-data-(\d+)-pipeline-(\d+)-tensor-(\d+)\.json$")
_COLOURS = {"forward": "thread_state_running", "backward": "thread_state_iowait", "optimizer": "rail_idle",
            "_reduce": "rail_response", "allreduce": "rail_response", "loss": "rail_load",
            "send-forward": "thread_state_runnable", "recv-forward": "thread_state_runnable",
            "send-backward": "thread_state_runnable", "recv-backward": "thread_state_runnable"}


def read_benchmark_dir(bench_dir: str) -> Dict[Tuple[int, int, int], List[dict]]:
    out = {}
    for path in sorted(glob.glob(os.path.join(bench_dir, "benchmark-data-*.json"))):
        m = _FILE_RE.search(os.path.basename(path))
        if not m:
            continue
        with open(path) as f:
            recs = json.load(f)
        out[tuple(int(x) for x in m.groups())] = recs
    for path in sorted(glob.glob(os.path.join(bench_dir, "benchmark-data-*.jsonl"))):
        m = _FILE_RE.search(os.path.basename(path)[:-1])
        if m:
            with open(path) as f:
                out.setdefault(tuple(int(x) for x in m.groups()), []).extend(json.loads(l) for l in f if l.strip())
    return out


def split_iterations(recs: List[dict]) -> List[List[dict]]:
    its, cur = [], None
    for r in recs:
        if r["name"] == "iteration" and r["ph"] == "B":
            cur = [r]
        elif cur is not None:
            cur.append(r)
            if r["name"] == "iteration" and r["ph"] == "E":
                its.append(cur)
                cur = None
    return its


def to_complete_events(it: List[dict], t0_ns: float) -> List[dict]:
    """Pair B/E records of one iteration into Chrome ``X`` events (ts/dur in microseconds)."""
    stack: List[dict] = []
    out = []
    for r in it:
        if r["ph"] == "B":
            stack.append(r)
        elif r["ph"] == "E":
            for k in range(len(stack) - 1, -1, -1):
                if stack[k]["name"] == r["name"]:
                    b = stack.pop(k)
                    args = {k2: v for k2, v in b.items() if k2 not in ("name", "ph", "rel_ts")}
                    args.update({k2: v for k2, v in r.items() if k2 in ("duration_wall", "duration_cuda")})
                    out.append({"name": b["name"], "ph": "X", "ts": (t0_ns + b["rel_ts"]) / 1e3,
                                "dur": max(0.0, (r["rel_ts"] - b["rel_ts"]) / 1e3), "args": args})
                    break
        elif r["ph"] == "i":
            out.append({"name": r["name"], "ph": "i", "ts": (t0_ns + r["rel_ts"]) / 1e3, "s": "t",
                        "args": {k2: v for k2, v in r.items() if k2 not in ("name", "ph", "rel_ts")}})
    return out


def benchmark_to_chrome_trace(data: Dict[Tuple[int, int, int], List[dict]], align: str = "iteration") -> List[dict]:
    """``align``: "iteration" stacks iterations back to back per rank from 0 (reference
    behaviour); "wall" uses the tracer's ``unix_ns`` anchors for a true cross-rank timeline."""
    traces: List[dict] = []
    anchors = []
    for recs in data.values():
        for it in split_iterations(recs):
            if "unix_ns" in it[0]:
                anchors.append(it[0]["unix_ns"])
    base = min(anchors) if anchors else 0
    for (d, p, t), recs in sorted(data.items()):
        g_rk = None
        cursor = 0.0
        for it in split_iterations(recs):
            g_rk = it[0].get("g_rk", 0)
            if align == "wall" and "unix_ns" in it[0]:
                t0 = float(it[0]["unix_ns"] - base)
            else:
                t0 = cursor + float(it[0].get("pad_before", 0) if cursor else 0)
            evs = to_complete_events(it, t0)
            for e in evs:
                e["pid"], e["tid"] = g_rk, 0
                e["cname"] = _COLOURS.get(e["name"], "generic_work")
            traces.extend(evs)
            cursor = t0 + it[-1]["rel_ts"]
        if g_rk is not None:
            traces.append({"name": "process_name", "ph": "M", "pid": g_rk, "args": {"name": f"DP{d}-PP{p}-TP{t}"}})
            traces.append({"name": "process_sort_index", "ph": "M", "pid": g_rk, "args": {"sort_index": g_rk}})
    traces.sort(key=lambda e: (e.get("ts", -1), e["pid"]))
    for i, e in enumerate(traces):
        e.setdefault("args", {})["id"] = i
    return traces


def dependency(traces: List[dict]) -> None:
    """Link events of the same collective: same (expect-name, sorted rank group), k-th occurrence
    on every member rank -> ``args.related_sync_op`` = space-separated event ids."""
    per_rank_counter: Dict[Tuple, int] = collections.Counter()
    buckets: Dict[Tuple, List[int]] = collections.defaultdict(list)
    for i, e in enumerate(traces):
        a = e.get("args", {})
        if e.get("ph") != "X" or "group" not in a or a.get("expect") is None:
            continue
        grp = tuple(sorted(a["group"]))
        key_rank = (a["expect"], grp, e["pid"])
        k = per_rank_counter[key_rank]
        per_rank_counter[key_rank] += 1
        buckets[(a["expect"], grp, k)].append(i)
    for ids in buckets.values():
        s = " ".join(str(traces[i]["args"]["id"]) for i in ids)
        for i in ids:
            traces[i]["args"]["related_sync_op"] = s


def amend_p2p(traces: List[dict]) -> None:
    """A send and its matching recv overlap only for the shorter of the two durations: trim both
    to it (and keep the larger bandwidth), so waiting time does not read as transfer time."""
    sends: Dict[Tuple, List[int]] = collections.defaultdict(list)
    recvs: Dict[Tuple, List[int]] = collections.defaultdict(list)
    for i, e in enumerate(traces):
        if e.get("ph") != "X":
            continue
        n, a = e["name"], e.get("args", {})
        if "peer" not in a:
            continue
        if n.startswith("send"):
            sends[(e["pid"], a["peer"], n[5:])].append(i)
        elif n.startswith("recv"):
            recvs[(a["peer"], e["pid"], n[5:])].append(i)
    for key, s_ids in sends.items():
        for si, ri in zip(s_ids, recvs.get(key, [])):
            s, r = traces[si], traces[ri]
            dur = min(s["dur"], r["dur"])
            end = min(s["ts"] + s["dur"], r["ts"] + r["dur"])
            for e in (s, r):
                e["args"]["orig_dur"] = e["dur"]
                e["ts"], e["dur"] = end - dur, dur
                if "data" in e["args"] and dur > 0:
                    e["args"]["bandwidth"] = e["args"]["data"] * 8 / (dur * 1e3)


# ---------------------------------------------------------------------------
# slow-rank detection
# ---------------------------------------------------------------------------

_SYNC_NAMES = {"_reduce", "_gather_along_first_dim", "_reduce_scatter_along_first_dim", "_gather_along_last_dim",
               "_reduce_scatter_along_last_dim", "ag_gemm", "gemm_rs", "gemm_rs_dgrad", "ag_gemm_dgrad", "ag_wgrad", "all_to_all"}


def _per_rank_durations(traces: List[dict], names) -> Dict[int, Dict[str, List[float]]]:
    out: Dict[int, Dict[str, List[float]]] = collections.defaultdict(lambda: collections.defaultdict(list))
    for e in traces:
        if e.get("ph") == "X" and e["name"] in names:
            out[e["pid"]][e["name"]].append(e["dur"])
    return out


def try_detect(traces: List[dict], out_file: Optional[str] = "abnormal.txt", compute_thresh: float = 1.1,
               wait_thresh: float = 0.9, min_votes: int = 3, sync_frac: float = 0.4) -> List[dict]:
    """Two-stage heuristic (reference :399-489, generalised):

    stage 1  among ranks that run the *same* work (same pp & tp coordinates, different dp; and
             same dp & pp, different tp) a rank whose compute scopes (``backward``, ``forward``,
             ``transformer_layer``) take > 1.1x the mean of its peers, or whose waiting scopes
             (``loss`` / ``allreduce``) take < 0.9x, collects a vote per event;
    stage 2  a rank with enough votes is confirmed if, in the collectives it took part in
             (``related_sync_op`` sets of ``_reduce`` & friends), it is the *fast* one
             (< 0.9x the peers' mean: it arrived last so it waited least) in > 40% of cases -- or, time-weighted, if
             its *total* time inside those collectives is < 0.9x the mean total of its peers (a rank that is late only
             at a few synchronisation points per step still makes every peer wait there: counting events misses it,
             summing the waits does not).
    """
    meta: Dict[int, Tuple[int, int, int]] = {}
    for e in traces:
        a = e.get("args", {})
        if "dp_rk" in a and e["pid"] not in meta:
            meta[e["pid"]] = (a["dp_rk"], a["pp_rk"], a["tp_rk"])
    compute_names = ("backward", "forward", "transformer_layer")
    wait_names = ("loss", "allreduce")
    durs = _per_rank_durations(traces, compute_names + wait_names)
    votes: Dict[int, int] = collections.Counter()

    def compare(group: List[int]):
        if len(group) < 2:
            return
        for name in compute_names + wait_names:
            series = [durs[r][name] for r in group if durs[r][name]]
            if len(series) < 2:
                continue
            n = min(len(s) for s in series)
            ranks = [r for r in group if durs[r][name]]
            for i in range(n):
                vals = [(durs[r][name][i], r) for r in ranks]
                vals.sort()
                if name in wait_names:
                    v, r = vals[0]
                    rest = np.mean([x[0] for x in vals[1:]])
                    if v < wait_thresh * rest:
                        votes[r] += 1
                else:
                    v, r = vals[-1]
                    rest = np.mean([x[0] for x in vals[:-1]])
                    if v > compute_thresh * rest:
                        votes[r] += 1

    by_pt, by_dp = collections.defaultdict(list), collections.defaultdict(list)
    for r, (d, p, t) in meta.items():
        by_pt[(p, t)].append(r)
        by_dp[(d, p)].append(r)
    for g in by_pt.values():
        compare(sorted(g))
    for g in by_dp.values():
        compare(sorted(g))
    # stage 1b: tensor-parallel peers run in lock step -- their collectives equalise every scope's duration, so a slow GPU
    # shows up only in *where* the time goes: its scopes hold more compute and less waiting.  Exclusive compute of the k-th
    # ``forward`` / ``backward`` / ``iteration`` scope = duration - time inside the collectives it contains.
    sync_ev: Dict[int, List[Tuple[float, float]]] = collections.defaultdict(list)
    for e in traces:
        if e.get("ph") == "X" and e["name"] in _SYNC_NAMES and "group" in e.get("args", {}):
            sync_ev[e["pid"]].append((e["ts"], e["dur"]))
    excl: Dict[int, Dict[str, List[float]]] = collections.defaultdict(lambda: collections.defaultdict(list))
    for r, evs in sync_ev.items():
        evs.sort()
        starts = np.array([t for t, _ in evs])
        csum = np.concatenate([[0.0], np.cumsum([d for _, d in evs])])
        for e in traces:
            if e.get("ph") == "X" and e["pid"] == r and e["name"] in ("forward", "backward", "iteration"):
                lo, hi = np.searchsorted(starts, e["ts"]), np.searchsorted(starts, e["ts"] + e["dur"])
                excl[r][e["name"]].append(e["dur"] - float(csum[hi] - csum[lo]))
    for g in by_dp.values():
        g = [r for r in sorted(g) if r in excl]
        for name in ("forward", "backward", "iteration"):
            series = {r: excl[r][name] for r in g if excl[r][name]}
            if len(series) < 2:
                continue
            for i in range(min(len(v) for v in series.values())):
                vals = sorted((v[i], r) for r, v in series.items())
                rest = float(np.mean([x[0] for x in vals[:-1]]))
                if rest > 0 and vals[-1][0] > compute_thresh * rest:
                    votes[vals[-1][1]] += 1
    findings = []
    by_id = {e["args"]["id"]: e for e in traces if "args" in e and "id" in e["args"]}
    sync_names = _SYNC_NAMES
    for r, v in votes.most_common():
        logging.info("Suspect: rank %s (%s) %d votes", r, meta.get(r), v)
        if v < min_votes:
            continue
        all_cnt = fast_cnt = 0
        own_total = peer_total = 0.0
        for e in traces:
            if e.get("ph") != "X" or e["pid"] != r or e["name"] not in sync_names:
                continue
            rel = e.get("args", {}).get("related_sync_op")
            if not rel:
                continue
            peers = [by_id[int(i)] for i in rel.split() if int(i) != e["args"]["id"] and int(i) in by_id]
            if not peers:
                continue
            all_cnt += 1
            peer_mean = float(np.mean([p["dur"] for p in peers]))
            own_total += e["dur"]
            peer_total += peer_mean
            if e["dur"] < wait_thresh * peer_mean:
                fast_cnt += 1
        confirmed = all_cnt == 0 or fast_cnt > sync_frac * all_cnt or own_total < wait_thresh * peer_total
        findings.append({"g_rk": r, "rank": meta.get(r), "votes": v, "sync_all": all_cnt, "sync_fast": fast_cnt,
                         "sync_time_ratio": round(own_total / peer_total, 4) if peer_total else None,
                         "confirmed": bool(confirmed)})
    confirmed = [f for f in findings if f["confirmed"]]
    if confirmed and out_file:
        with open(out_file, "w") as f:
            for c in confirmed[:1]:
                d, p, t = c["rank"]
                f.write(f"{c['g_rk']} {d}-{p}-{t}\n")
                logging.info("Abnormal GPU: %s %s", c["g_rk"], c["rank"])
    return findings


def aggregate(bench_dir: str, output: str, detect: bool = False, align: str = "iteration",
              abnormal_file: str = "abnormal.txt"):
    data = read_benchmark_dir(bench_dir)
    if not data:
        raise FileNotFoundError(f"no benchmark-data-*.json under {bench_dir}")
    traces = benchmark_to_chrome_trace(data, align)
    dependency(traces)
    amend_p2p(traces)
    with open(output, "w") as f:
        json.dump(traces, f)
    findings = try_detect(traces, abnormal_file) if detect else []
    return traces, findings


def main(argv=None):
    ap = argparse.ArgumentParser(description="MegaScan trace aggregator / slow-rank detector")
    ap.add_argument("-b", "--bench-dir", required=True)
    ap.add_argument("-o", "--output", default="aggregated.json")
    ap.add_argument("-d", "--detect", action="store_true")
    ap.add_argument("--align", default="iteration", choices=["iteration", "wall"])
    ap.add_argument("--abnormal-file", default="abnormal.txt")
    a = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    traces, findings = aggregate(a.bench_dir, a.output, a.detect, a.align, a.abnormal_file)
    logging.info("wrote %d events to %s", len(traces), a.output)
    for f in findings:
        logging.info("finding: %s", f)


if __name__ == "__main__":
    main()
