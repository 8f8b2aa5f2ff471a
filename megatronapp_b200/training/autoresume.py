"""Cluster auto-resume hook (reference ``--adlr-autoresume``: ``global_vars._set_adlr_autoresume`` +
``training/utils.py check_adlr_autoresume_termination``).

The reference imports NVIDIA's cluster-internal ``userlib.auto_resume.AutoResume``.  Here that library is used when
importable; otherwise a file / deadline protocol gives the same behaviour on any scheduler: termination is
"requested" when the sentinel ``$B200_AUTORESUME_FILE`` exists or the wall-clock deadline
``$B200_AUTORESUME_DEADLINE`` (unix seconds, e.g. the SLURM end time minus a margin) has passed;
``request_resume()`` leaves ``<sentinel>.resume`` for the launcher's requeue hook.
"""
from __future__ import annotations

import os
import sys
import time
from typing import Optional

import torch

_AUTORESUME = None


class FileAutoResume:
    def __init__(self, sentinel: Optional[str] = None, deadline: Optional[float] = None):
        self.sentinel = sentinel if sentinel is not None else os.environ.get("B200_AUTORESUME_FILE")
        d = deadline if deadline is not None else os.environ.get("B200_AUTORESUME_DEADLINE")
        self.deadline = float(d) if d else None
        self.resume_requested = False

    def init(self):
        pass

    def termination_requested(self) -> bool:
        if self.sentinel and os.path.exists(self.sentinel):
            return True
        return self.deadline is not None and time.time() >= self.deadline

    def request_resume(self):
        self.resume_requested = True
        if self.sentinel:
            with open(self.sentinel + ".resume", "w") as f:
                f.write(str(time.time()))


def build_autoresume():
    global _AUTORESUME
    try:
        sys.path.append(os.environ.get("SUBMIT_SCRIPTS", "."))
        from userlib.auto_resume import AutoResume  # type: ignore

        _AUTORESUME = AutoResume
    except ImportError:
        _AUTORESUME = FileAutoResume()
    _AUTORESUME.init()
    return _AUTORESUME


def get_adlr_autoresume():
    return _AUTORESUME


def set_adlr_autoresume(obj):
    global _AUTORESUME
    _AUTORESUME = obj


def check_adlr_autoresume_termination(iteration, model, optimizer, opt_param_scheduler, save_fn=None) -> bool:
    """All ranks agree (MAX all-reduce of rank-local polls), save, rank 0 asks for a requeue.  Returns True when the
    training loop must stop (the reference calls ``sys.exit(0)`` here; the caller decides)."""
    ar = get_adlr_autoresume()
    if ar is None:
        return False
    dist = torch.distributed
    flag = int(bool(ar.termination_requested()))
    if dist.is_available() and dist.is_initialized():
        dev = "cuda" if (torch.cuda.is_available() and dist.get_backend() == "nccl") else "cpu"
        t = torch.tensor([flag], dtype=torch.int, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        flag = int(t.item())
    if not flag:
        return False
    if save_fn is not None:
        save_fn(iteration, model, optimizer, opt_param_scheduler)
    rank0 = not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
    if rank0:
        print(">>> autoresume termination request found", flush=True)
        ar.request_resume()
        print(">>> training terminated, resume requested", flush=True)
    return True
