"""Named hierarchical timers (reference ``core/timers.py``: log levels 0-2, optional barrier,
min/max/all aggregation across ranks).  Device time is taken from CUDA events when possible so a
``stop()`` does not force ``cuda.synchronize`` on the hot path; ``elapsed`` synchronises once."""
from __future__ import annotations

import time
from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class _Timer:
    def __init__(self, name):
        self.name = name
        self._elapsed = 0.0
        self._started = False
        self._pairs = []
        self._t0 = None
        self._cuda = torch.cuda.is_available()
        self._active_time = 0.0
        self._count = 0

    def start(self, barrier=False):
        if self._started:
            raise RuntimeError(f"timer {self.name} already started")
        if barrier and dist.is_initialized():
            dist.barrier()
        if self._cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._t0 = ev
        else:
            self._t0 = time.perf_counter()
        self._started = True

    def stop(self, barrier=False):
        if not self._started:
            raise RuntimeError(f"timer {self.name} is not started")
        if barrier and dist.is_initialized():
            dist.barrier()
        if self._cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._pairs.append((self._t0, ev))
        else:
            self._elapsed += time.perf_counter() - self._t0
        self._started = False
        self._count += 1

    def reset(self):
        self._elapsed = 0.0
        self._pairs = []
        self._started = False

    def elapsed(self, reset=True, barrier=False):
        was = self._started
        if was:
            self.stop(barrier)
        if self._pairs:
            torch.cuda.synchronize()
            for a, b in self._pairs:
                self._elapsed += a.elapsed_time(b) / 1000.0
            self._pairs = []
        e = self._elapsed
        self._active_time += e if reset else 0.0
        if reset:
            self._elapsed = 0.0
        if was:
            self.start(barrier)
        return e

    def active_time(self):
        return self._active_time


class _Dummy:
    def start(self, barrier=False):
        pass

    def stop(self, barrier=False):
        pass

    def reset(self):
        pass

    def elapsed(self, reset=True, barrier=False):
        return 0.0

    def active_time(self):
        return 0.0


class Timers:
    def __init__(self, log_level: int = 0, log_option: str = "minmax"):
        self._log_level, self._log_option = log_level, log_option
        self._timers: Dict[str, _Timer] = {}
        self._levels: Dict[str, int] = {}
        self._dummy = _Dummy()

    def __call__(self, name, log_level=None):
        if name in self._timers:
            return self._timers[name]
        lvl = 2 if log_level is None else log_level
        if lvl > self._log_level:
            return self._dummy
        self._timers[name] = _Timer(name)
        self._levels[name] = lvl
        return self._timers[name]

    def get_all_timers_string(self, names: Optional[List[str]] = None, normalizer=1.0, reset=True, barrier=False):
        names = names or list(self._timers)
        vals = {n: self._timers[n].elapsed(reset, barrier) * 1000.0 / normalizer for n in names if n in self._timers}
        if not vals:
            return ""
        if dist.is_initialized() and dist.get_world_size() > 1 and self._log_option in ("minmax", "max", "all"):
            dev = "cuda" if (torch.cuda.is_available() and dist.get_backend() == "nccl") else "cpu"
            t = torch.tensor(list(vals.values()), dtype=torch.float32, device=dev)
            mx, mn = t.clone(), t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            if self._log_option == "max":
                return "max time across ranks (ms):" + "".join(
                    f"\n    {n}: {float(v):.2f}" for n, v in zip(vals, mx.tolist()))
            return "(min, max) time across ranks (ms):" + "".join(
                f"\n    {n}: ({float(a):.2f}, {float(b):.2f})" for n, a, b in zip(vals, mn.tolist(), mx.tolist()))
        return "time (ms):" + "".join(f"\n    {n}: {v:.2f}" for n, v in vals.items())

    def log(self, names=None, rank=None, normalizer=1.0, reset=True, barrier=False):
        s = self.get_all_timers_string(names, normalizer, reset, barrier)
        if not s:
            return
        r = dist.get_rank() if dist.is_initialized() else 0
        want = (dist.get_world_size() - 1) if (rank is None and dist.is_initialized()) else (rank or 0)
        if r == want:
            print(s, flush=True)

    def write(self, names, writer, iteration, normalizer=1.0, reset=True, barrier=False):
        for n in names:
            if n in self._timers:
                writer.add_scalar(n + "-time", self._timers[n].elapsed(reset, barrier) / normalizer, iteration)


TimerBase = _Timer          # class names of the reference's ``core/timers.py``
Timer = _Timer
DummyTimer = _Dummy
