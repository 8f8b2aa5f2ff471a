"""Fault-tolerance hooks (reference ``training/ft_integration.py`` -- NVRx rank-monitor client):
section heartbeats (setup / step / checkpointing) with timeouts.  Without the external
``nvidia_resiliency_ext`` launcher the client degrades to a local watchdog thread that dumps stacks
and aborts the rank when a section exceeds its timeout, so hangs surface instead of idling."""
from __future__ import annotations

import faulthandler
import os
import sys
import threading
import time
from typing import Dict, Optional

_STATE: Dict[str, object] = {"enabled": False, "section": None, "deadline": None, "timeouts": {}, "thread": None}
_DEFAULT_TIMEOUTS = {"setup": 1800.0, "step": 600.0, "checkpointing": 1800.0, "out_of_section": 3600.0}


def setup(args=None, timeouts: Optional[Dict[str, float]] = None) -> None:
    _STATE["timeouts"] = {**_DEFAULT_TIMEOUTS, **(timeouts or {})}
    _STATE["enabled"] = True
    if _STATE["thread"] is None:
        t = threading.Thread(target=_watch, daemon=True)
        _STATE["thread"] = t
        t.start()
    on_training_step_end()


def _watch():
    while _STATE["enabled"]:
        d = _STATE["deadline"]
        if d is not None and time.time() > d:
            sys.stderr.write(f"[ft] section {_STATE['section']!r} exceeded its timeout; dumping stacks and aborting\n")
            faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
            os._exit(int(os.environ.get("B200_FT_EXIT_CODE", "18")))
        time.sleep(1.0)


def _enter(section: str):
    if _STATE["enabled"]:
        _STATE["section"] = section
        _STATE["deadline"] = time.time() + _STATE["timeouts"].get(section, 3600.0)


def on_training_step_start():
    _enter("step")


def on_training_step_end():
    _enter("out_of_section")


def on_eval_step_start():
    _enter("step")


on_eval_step_end = on_training_step_end


def on_checkpointing_start():
    _enter("checkpointing")


on_checkpointing_end = on_training_step_end


def on_checkpoint_loaded(is_local_chkpt: bool = False):
    on_training_step_end()


def shutdown():
    _STATE["enabled"] = False
    _STATE["deadline"] = None


def get_rank_monitor_client():
    return None
