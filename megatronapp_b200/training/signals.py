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


# -- helper names of the reference's ``dist_signal_handler.py`` --------------------------------------------------------------------------
def get_world_size() -> int:
    import torch.distributed as dist

    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def get_device(local_rank=None):
    """Device the signal flags are gathered on: the rank's GPU under NCCL, the CPU under gloo."""
    import torch
    import torch.distributed as dist

    backend = dist.get_backend() if dist.is_available() and dist.is_initialized() else "gloo"
    if backend == "nccl":
        return torch.device("cuda", torch.cuda.current_device() if local_rank is None else local_rank)
    return torch.device("cpu")


def all_gather_item(item, dtype, group=None, async_op: bool = False, local_rank=None):
    """Every rank's scalar ``item`` as a python list (one entry per rank of ``group``)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return [item]
    dev = get_device(local_rank)
    group = group if group is not None else dist.group.WORLD
    t = torch.tensor([item], device=dev, dtype=dtype)
    outs = [torch.zeros(1, dtype=t.dtype, device=dev) for _ in range(dist.get_world_size(group))]
    dist.all_gather(outs, t, group, async_op)
    return [o.item() for o in outs]
