"""SIGTERM -> checkpoint-and-exit across all ranks (reference ``training/dist_signal_handler.py``)."""
from __future__ import annotations

import signal

import torch
import torch.distributed as dist


class DistributedSignalHandler:
    def __init__(self, sig=signal.SIGTERM):
        self.sig = sig
        self._received = False
        self.released = False
        self.original = None

    def signals_received(self):
        if not dist.is_initialized():
            return [self._received]
        dev = "cuda" if (torch.cuda.is_available() and dist.get_backend() == "nccl") else "cpu"
        t = torch.tensor([int(self._received)], dtype=torch.int32, device=dev)
        out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        return [bool(x.item()) for x in out]

    def __enter__(self):
        self._received = False
        self.released = False
        self.original = signal.getsignal(self.sig)

        def handler(signum, frame):
            self._received = True

        signal.signal(self.sig, handler)
        return self

    def __exit__(self, *a):
        self.release()

    def release(self):
        if self.released:
            return False
        signal.signal(self.sig, self.original)
        self.released = True
        return True
