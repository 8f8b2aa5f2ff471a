"""MegaScan tracer: scoped, sync-free timing of every rank.

``scope(name, **attrs)`` records a begin/end pair per call site.  On CUDA the pair is two
``torch.cuda.Event(enable_timing=True)`` records on the current stream (no host sync on the hot
path); on CPU (gloo plumbing runs) it is ``time.perf_counter_ns``.  ``iteration_end`` does NOT
synchronise the device: it records the end event and hands the iteration's scope tree to a resolver
thread, which waits on that event (a host-side wait, the GPU keeps running the next step), converts
event pairs to nanosecond offsets from the iteration-begin event and recycles the events through a pool
(no ``cudaEventCreate`` on the hot path after the first traced iteration).  With the in-tree extension the events live in a
native pool (``_C.trace_tick`` -> an int; one ``_C.trace_resolve`` call per iteration converts them all with the GIL
released).  Records are tagged with dp/pp/tp rank, device and global rank and streamed to disk *by the rank that produced
them* (background thread, JSON-lines, finalised to one JSON array per rank-tuple file on shutdown): on one NVSwitch node
every rank sees the same trace directory, so nothing is gathered on the training path; ranks on other nodes hand their
files to rank 0 once, at shutdown.

Capability parity: reference ``megatron/training/trace.py`` (``Tracer`` :242, ``_TracerScope``
:196, windows :594-614, granularity :75-100, file name :544, bandwidth :371-380).
Deliberate fixes: no WORLD/TP barriers, ``iteration_end`` is idempotent,
pendings are cleared per iteration, append-only I/O, and an optional cross-rank clock offset
(``unix_ns`` of the iteration begin) so the aggregator can align timelines.
"""
from __future__ import annotations

import contextlib
import json
import os
import queue
import threading
import time
from typing import Any, Callable, Dict, List, Optional

import torch

BASE_TRACING_EVENTS = {
    "iteration", "forward", "backward", "optimizer", "loss", "allreduce", "reduce",
    "send-forward", "recv-forward", "send-backward", "recv-backward",
    "send-forward-recv-backward", "send-backward-recv-forward", "decoder", "transformer_layer",
    "_reduce", "_gather_along_first_dim", "_reduce_scatter_along_first_dim",
    "_gather_along_last_dim", "_reduce_scatter_along_last_dim", "all_to_all",
    "grad-reduce-scatter", "param-all-gather", "fbd-send", "fbd-recv",
}

_NULL = contextlib.nullcontext()


class _Scope:
    __slots__ = ("tracer", "name", "attrs", "b", "e", "children", "parent", "slots")

    def __init__(self, tracer: "Tracer", name: str, attrs: Dict[str, Any]):
        self.tracer, self.name, self.attrs = tracer, name, attrs
        self.b = self.e = None
        self.children: List["_Scope"] = []
        self.parent: Optional["_Scope"] = None
        self.slots: Dict[str, Any] = {}

    def __enter__(self):
        t = self.tracer
        self.parent = t._stack[-1] if t._stack else None
        (self.parent.children if self.parent is not None else t._pending).append(self)
        t._stack.append(self)
        self.b = t._tick()
        return self

    def __exit__(self, *exc):
        self.e = self.tracer._tick()
        self.tracer._stack.pop()
        return False

    def set(self, key: str, value: Any):
        self.slots[key] = value


class Tracer:
    def __init__(self):
        self.enabled = False
        self.trace_dir = "trace"
        self.interval = 1
        self.continuous = 1
        self.granularity = "full"
        self.active = False
        self.iteration = -1
        self._pending: List[_Scope] = []
        self._stack: List[_Scope] = []
        self._records: List[dict] = []
        self._iter_begin = None
        self._iter_begin_wall = 0
        self._prev_end_wall = 0
        self._ended = True
        self._use_cuda = False
        self._rank_info: Dict[str, int] = {}
        self._q: Optional[queue.Queue] = None
        self._thread: Optional[threading.Thread] = None
        self._files: Dict[str, str] = {}
        # sync-free resolution: finished iterations wait here for their end event
        self._rq: Optional[queue.Queue] = None
        self._rthread: Optional[threading.Thread] = None
        self._free: List[Any] = []            # recycled CUDA events (list.append / list.pop are atomic)
        self._ship_at: Optional[int] = None
        self._device = None
        self._native = None                   # _C module when its event pool is usable
        self._switch_interval: Optional[float] = None

    # -- configuration -----------------------------------------------------------
    def configure(self, enabled: bool, trace_dir: str = "trace", interval: int = 1,
                  continuous_iterations: int = 1, granularity: str = "full") -> None:
        self.enabled = enabled
        self.trace_dir = trace_dir
        self.interval = max(1, interval)
        self.continuous = max(1, continuous_iterations)
        self.granularity = granularity
        self._use_cuda = torch.cuda.is_available()
        self._native = None
        if self._use_cuda and enabled:
            try:
                from .. import _C

                self._native = _C if hasattr(_C, "trace_tick") else None
            except ImportError:
                self._native = None
        if enabled and self._switch_interval is None:
            # the resolver / writer threads are pure Python: a short GIL switch interval keeps the launching thread from
            # waiting the default 5 ms behind them
            import sys

            self._switch_interval = sys.getswitchinterval()
            sys.setswitchinterval(min(self._switch_interval, 2e-4))
        elif not enabled and self._switch_interval is not None:
            import sys

            sys.setswitchinterval(self._switch_interval)
            self._switch_interval = None

    def _ranks(self) -> Dict[str, int]:
        if not self._rank_info:
            from ..parallel import state as ps

            g_rk = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
            info = {"dp_rk": 0, "pp_rk": 0, "tp_rk": 0, "g_rk": g_rk,
                    "dev": torch.cuda.current_device() if self._use_cuda else -1}
            if ps.model_parallel_is_initialized():
                c = ps.get_state().coords
                info.update(dp_rk=c["dp"], pp_rk=c["pp"], tp_rk=c["tp"])
            self._rank_info = info
        return self._rank_info

    # -- hot path ----------------------------------------------------------------
    def _tick(self):
        if self._native is not None:
            return self._native.trace_tick()
        if self._use_cuda:
            try:
                ev = self._free.pop()
            except IndexError:
                ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            return ev
        return time.perf_counter_ns()

    def should_trace(self, iteration: int) -> bool:
        return self.enabled and (iteration % self.interval) < self.continuous

    def scope(self, name: str, **attrs):
        if not self.active:
            return _NULL
        if self.granularity == "base" and name not in BASE_TRACING_EVENTS:
            return _NULL
        return _Scope(self, name, attrs)

    def current(self) -> Optional[_Scope]:
        return self._stack[-1] if (self.active and self._stack) else None

    def set(self, key: str, value: Any) -> None:
        """Fill an attribute ("slot") of the innermost open scope from an inner call site."""
        s = self.current()
        if s is not None:
            s.set(key, value)

    def instant(self, name: str, **attrs) -> None:
        if not self.active:
            return
        s = _Scope(self, name, attrs)
        s.b = s.e = self._tick()
        (self._stack[-1].children if self._stack else self._pending).append(s)
        s.slots["__instant__"] = True

    # -- iteration window ----------------------------------------------------------
    def iteration_begin(self, iteration: int) -> None:
        self.iteration = iteration
        self.active = self.should_trace(iteration)
        self._ended = False
        if not self.active:
            return
        self._pending, self._stack = [], []
        now = time.time_ns()
        self._pad_before = now - self._prev_end_wall if self._prev_end_wall else 0
        self._iter_begin_wall = now
        self._iter_begin = self._tick()

    def iteration_end(self) -> None:
        if self._ended:
            return
        self._ended = True
        if self._ship_at is not None and self.iteration >= self._ship_at:
            # the window that closed one iteration ago has certainly left the GPU by now: shipping it does not stall
            self._ship_at = None
            self.log()
        if not self.active:
            return
        end = self._tick()
        job = {"iteration": self.iteration, "begin": self._iter_begin, "end": end, "pending": self._pending,
               "pad_before": self._pad_before, "wall_begin": self._iter_begin_wall, "info": self._ranks()}
        self._pending = []
        self._prev_end_wall = time.time_ns()
        self.active = False
        if self._use_cuda:
            if self._rthread is None:
                self._device = torch.cuda.current_device()
                self._rq = queue.Queue()
                self._rthread = threading.Thread(target=self._resolver, name="megascan-resolver", daemon=True)
                self._rthread.start()
            self._rq.put(job)
        else:
            self._resolve(job)
        if (self.iteration % self.interval) == self.continuous - 1:
            self._ship_at = self.iteration + 1      # ship one iteration later (see iteration_end); shutdown() flushes

    def _resolver(self) -> None:
        torch.cuda.set_device(self._device)
        while True:
            job = self._rq.get()
            try:
                if job is not None:
                    # host-side wait; the training stream is not stalled
                    self._native.trace_wait(job["end"]) if self._native is not None else job["end"].synchronize()
                    self._resolve(job)
            finally:
                self._rq.task_done()
            if job is None:
                return

    def _drain(self) -> None:
        """Wait until every finished iteration has been converted to records (called before shipping them)."""
        if self._rq is not None:
            self._rq.join()

    def _resolve(self, job) -> None:
        begin, end, info = job["begin"], job["end"], job["info"]
        wall_end = time.time_ns()
        if self._native is not None:
            # one GIL-free call turns every event of the iteration into a nanosecond offset (and recycles the events)
            idx: List[int] = [end]
            stack = list(job["pending"])
            while stack:
                sc = stack.pop()
                if sc.e is None:
                    continue
                idx.append(sc.b)
                if sc.e != sc.b:
                    idx.append(sc.e)
                stack.extend(sc.children)
            table = dict(zip(idx, self._native.trace_resolve(begin, idx, True)))
            rel = table.__getitem__
        elif self._use_cuda:
            rel = lambda t: int(begin.elapsed_time(t) * 1e6)  # noqa: E731
        else:
            rel = lambda t: int(t - begin)  # noqa: E731
        recs = [{"name": "iteration", "ph": "B", "rel_ts": 0, "iteration": job["iteration"],
                 "pad_before": job["pad_before"], "unix_ns": job["wall_begin"], **info}]
        used: List[Any] = [begin, end]
        for s in job["pending"]:
            self._emit(s, recs, info, rel, used)
        dur = rel(end)
        recs.append({"name": "iteration", "ph": "E", "rel_ts": dur, "iteration": job["iteration"],
                     "duration_wall": wall_end - job["wall_begin"], "duration_cuda": dur, **info})
        self._records.extend(recs)
        if self._use_cuda and self._native is None:
            self._free.extend(used)

    def _emit(self, s: _Scope, out: List[dict], info: Dict[str, int], rel, used: List[Any]) -> None:
        if s.e is None:
            return
        attrs = {}
        for k, v in {**s.attrs, **s.slots}.items():
            if k == "__instant__":
                continue
            if callable(v):
                try:
                    v = v()
                except Exception:
                    v = None
            if v is not None:
                attrs[k] = v
        tb = rel(s.b)
        if s.slots.get("__instant__"):
            used.append(s.b)
            out.append({"name": s.name, "ph": "i", "rel_ts": tb, **attrs, **info})
            return
        te = rel(s.e)
        used.append(s.b)
        used.append(s.e)
        if "data" in attrs and te > tb:
            attrs["bandwidth"] = attrs["data"] * 8 / (te - tb)  # Gbit/s (bytes*8 / ns)
        out.append({"name": s.name, "ph": "B", "rel_ts": tb, **attrs, **info})
        for c in s.children:
            self._emit(c, out, info, rel, used)
        out.append({"name": s.name, "ph": "E", "rel_ts": te, **info})

    # -- output --------------------------------------------------------------------
    def file_name(self, info: Dict[str, int]) -> str:
        return f"benchmark-data-{info['dp_rk']}-pipeline-{info['pp_rk']}-tensor-{info['tp_rk']}.json"

    def log(self) -> None:
        """Hand this window's records to the local writer thread (no collective: see the module docstring)."""
        self._drain()
        recs, self._records = self._records, []
        if recs:
            self._ensure_writer()
            self._q.put(recs)

    def _ensure_writer(self) -> None:
        if self._thread is None:
            os.makedirs(self.trace_dir, exist_ok=True)
            self._q = queue.Queue()
            self._thread = threading.Thread(target=self._writer, name="megascan-writer", daemon=True)
            self._thread.start()

    def _writer(self) -> None:
        while True:
            item = self._q.get()
            if item is None:
                return
            info = item[0]
            os.makedirs(self.trace_dir, exist_ok=True)
            path = os.path.join(self.trace_dir, self.file_name(info) + "l")  # .jsonl sidecar
            self._files[path] = path[:-1]
            with open(path, "a") as f:
                for rec in item:
                    f.write(json.dumps(rec) + "\n")
                f.flush()
                os.fsync(f.fileno())

    def shutdown(self) -> None:
        self._drain()
        self._ship_at = None
        if self._records:
            self.log()
        if self._rthread is not None:
            self._rq.put(None)
            self._rthread.join()
            self._rthread = self._rq = None
        if self._thread is not None:
            self._q.put(None)
            self._thread.join()
            self._thread = None
            for jl, final in self._files.items():
                with open(jl) as f:
                    recs = [json.loads(line) for line in f if line.strip()]
                prev = []
                if os.path.exists(final):
                    try:
                        prev = json.load(open(final))
                    except Exception:
                        prev = []
                with open(final, "w") as f:
                    json.dump(prev + recs, f)
                os.remove(jl)
        self._collect_remote_nodes(list(self._files.values()))
        self._files = {}

    def _collect_remote_nodes(self, mine: List[str]) -> None:
        """Multi-node jobs: ranks that do not share rank 0's trace directory send their finished files to rank 0.  Runs once,
        in ``shutdown`` (collective over WORLD when tracing was enabled on every rank), never on the training path."""
        if not (self.enabled and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1):
            return
        import socket

        world, rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
        if int(os.environ.get("LOCAL_WORLD_SIZE", world)) >= world:
            return                                       # one node: every rank already wrote into the shared directory
        hosts = [None] * world
        try:
            torch.distributed.all_gather_object(hosts, socket.gethostname())
        except Exception:
            return
        if len(set(hosts)) == 1:
            return
        payload = None
        if hosts[rank] != hosts[0]:
            payload = {os.path.basename(f): open(f).read() for f in mine if os.path.exists(f)}
        got = [None] * world if rank == 0 else None
        torch.distributed.gather_object(payload, got, dst=0)
        if rank == 0:
            os.makedirs(self.trace_dir, exist_ok=True)
            for files in got:
                for name, text in (files or {}).items():
                    path = os.path.join(self.trace_dir, name)
                    prev = json.load(open(path)) if os.path.exists(path) else []
                    with open(path, "w") as f:
                        json.dump(prev + json.loads(text), f)


_TRACER = Tracer()


def get_tracer() -> Tracer:
    return _TRACER


def scope(name: str, **attrs):
    """Module-level shortcut; returns a shared null context when tracing is inactive."""
    t = _TRACER
    if not t.active:
        return _NULL
    return t.scope(name, **attrs)
