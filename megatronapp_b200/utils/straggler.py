"""Online straggler detection (reference ``core/utils.py StragglerDetector :1030-1500``): per-rank
CUDA-event + wall timing around batch fetch and forward/backward, periodic all-gather and a
min/max report (time, TFLOP/s, which rank), runtime on/off through a tiny TCP control port."""
from __future__ import annotations

import socket
import threading
import time
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist


@dataclass
class StragglerReport:
    min_elapsed: float = 0.0
    max_elapsed: float = 0.0
    min_rank: int = 0
    max_rank: int = 0
    min_tflops: float = 0.0
    max_tflops: float = 0.0
    slowest_ranks: Optional[List[int]] = None


class StragglerDetector:
    def __init__(self):
        self.enabled = False
        self.world, self.rank, self.mmcnt, self.port = 1, 0, 1, 65535
        self._bdata = False
        self._t0 = 0.0
        self._ev = None
        self.times: List[float] = []
        self.batch_times: List[float] = []
        self._srv: Optional[socket.socket] = None

    def configure(self, world: int, rank: int, mmcnt: int = 1, enabled: bool = True, port: int = 65535):
        self.world, self.rank, self.mmcnt, self.enabled, self.port = world, rank, mmcnt, enabled, port
        if rank == 0 and port and port < 65535:
            self._start_controller()

    def _start_controller(self):
        def serve():
            try:
                s = socket.socket()
                s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
                s.bind(("127.0.0.1", self.port))
                s.listen(1)
                self._srv = s
                while True:
                    c, _ = s.accept()
                    cmd = c.recv(64).decode().strip().lower()
                    if cmd in ("on", "off"):
                        self.enabled = cmd == "on"
                        c.send(f"straggler detector {cmd}\n".encode())
                    else:
                        c.send(f"enabled={self.enabled}\n".encode())
                    c.close()
            except OSError:
                pass

        threading.Thread(target=serve, daemon=True).start()

    def __call__(self, bdata: bool = False):
        self._bdata = bdata
        return self

    def __enter__(self):
        if not self.enabled:
            return self
        self._t0 = time.perf_counter()
        if torch.cuda.is_available() and not self._bdata:
            self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self._ev[0].record()
        return self

    def __exit__(self, *a):
        if not self.enabled:
            return False
        if self._bdata:
            self.batch_times.append(time.perf_counter() - self._t0)
        elif self._ev is not None:
            self._ev[1].record()
            self.times.append(self._ev)
        else:
            self.times.append(time.perf_counter() - self._t0)
        self._bdata = False
        return False

    def _drain(self) -> float:
        tot = 0.0
        for t in self.times:
            if isinstance(t, tuple):
                t[1].synchronize()
                tot += t[0].elapsed_time(t[1]) / 1e3
            else:
                tot += t
        self.times = []
        return tot

    def report(self, total_flops: float = 0.0, log_interval: int = 1) -> Optional[StragglerReport]:
        if not self.enabled:
            return None
        mine = self._drain() / max(1, log_interval)
        self.batch_times = []
        vals = [mine]
        if dist.is_initialized() and self.world > 1:
            dev = "cuda" if (torch.cuda.is_available() and dist.get_backend() == "nccl") else "cpu"
            t = torch.tensor([mine], dtype=torch.float64, device=dev)
            out = [torch.zeros_like(t) for _ in range(self.world)]
            dist.all_gather(out, t)
            vals = [float(x) for x in out]
        order = sorted(range(len(vals)), key=lambda i: vals[i])
        lo, hi = order[0], order[-1]
        fl = (total_flops / self.world) if total_flops else 0.0
        rep = StragglerReport(vals[lo], vals[hi], lo, hi,
                              fl / vals[hi] / 1e12 if vals[hi] > 0 else 0.0, fl / vals[lo] / 1e12 if vals[lo] > 0 else 0.0,
                              list(reversed(order[-self.mmcnt:])))
        if self.rank == 0:
            print(f" straggler report | min {rep.min_elapsed * 1e3:.1f} ms (rank {rep.min_rank}) | max "
                  f"{rep.max_elapsed * 1e3:.1f} ms (rank {rep.max_rank}) | slowest ranks {rep.slowest_ranks}", flush=True)
        return rep


_DETECTOR = StragglerDetector()


def get_straggler_detector() -> StragglerDetector:
    return _DETECTOR
