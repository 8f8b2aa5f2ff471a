"""Fault-tolerance integration: section heartbeats with timeouts, learned timeouts, simulated faults.

Parity: reference ``megatron/training/ft_integration.py`` (``--enable-ft-package``, ``--calc-ft-timeouts``; hooks
``setup`` :76, ``on_training_step_start/end`` :120-143, ``on_eval_step_start/end``, ``on_checkpointing_start/end`` :168-190,
``on_checkpoint_loaded`` :192, ``shutdown`` :206, ``maybe_setup_simulated_fault`` :299).  The reference delegates to the NVRx
``RankMonitorClient`` (an external launcher-side monitor); here the monitor is in-process: every rank runs a watchdog thread
that knows which section the rank is in (``setup`` | ``step`` | ``checkpointing`` | out of section) and when it must be left;
a rank that overstays dumps the stacks of all its threads and exits with ``B200_FT_EXIT_CODE`` (default 18), which takes the
job down through the launcher (torchrun kills the other ranks) so that it can be restarted from the last checkpoint instead
of idling inside a hung collective.

``--calc-ft-timeouts``: observed section durations (max over ranks) times ``SAFETY_FACTOR`` replace the default timeouts and
are written to ``<save>/ft_state.json``; the next run loads them.  As in the reference the ``setup`` timeout is only learned
once a persistent checkpoint was loaded (a cold start is not representative), ``step`` after ``MIN_ITERS`` iterations past the
warm-up ones, ``checkpointing`` only with synchronous saves, and the out-of-section timeout only from a run that saw everything.
"""
from __future__ import annotations

import faulthandler
import json
import os
import random
import signal
import sys
import threading
import time
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

SAFETY_FACTOR = 5.0
NUM_WARMUP_ITERS = 1                      # the first iteration (allocator warm-up, lazy init) is not timed as a step
MIN_ITERS_FOR_STEP_TIMEOUT_UPDATE = 16
_DEFAULT_TIMEOUTS = {"setup": 1800.0, "step": 600.0, "checkpointing": 1800.0, "out_of_section": 3600.0}


class RankMonitor:
    """Watchdog + bookkeeping of one rank (the in-process stand-in for NVRx's ``RankMonitorClient``)."""

    def __init__(self, timeouts: Optional[Dict[str, float]] = None, poll_s: float = 1.0):
        self.section_timeouts: Dict[str, float] = {**_DEFAULT_TIMEOUTS, **(timeouts or {})}
        self.section: Optional[str] = None
        self._entered = time.monotonic()
        self._deadline: Optional[float] = self._entered + self.section_timeouts["out_of_section"]
        self.observed: Dict[str, float] = {}              # longest stay per section (seconds)
        self._lock = threading.Lock()
        self._stop = threading.Event()
        self._poll = poll_s
        self._thread = threading.Thread(target=self._watch, name="ft-rank-monitor", daemon=True)
        self._thread.start()

    # -- sections ---------------------------------------------------------------------------
    def _record(self, now: float):
        name = self.section or "out_of_section"
        self.observed[name] = max(self.observed.get(name, 0.0), now - self._entered)

    def start_section(self, name: str):
        with self._lock:
            now = time.monotonic()
            self._record(now)
            self.section, self._entered = name, now
            self._deadline = now + self.section_timeouts.get(name, _DEFAULT_TIMEOUTS["out_of_section"])

    def end_section(self, name: Optional[str] = None):
        with self._lock:
            if name is not None and self.section != name:
                return
            now = time.monotonic()
            self._record(now)
            self.section, self._entered = None, now
            self._deadline = now + self.section_timeouts["out_of_section"]

    # -- watchdog -----------------------------------------------------------------------------
    def _watch(self):
        while not self._stop.wait(self._poll):
            with self._lock:
                late = self._deadline is not None and time.monotonic() > self._deadline
                name = self.section or "out_of_section"
            if late:
                sys.stderr.write(f"[ft] rank {_rank()}: section {name!r} exceeded its timeout of "
                                 f"{self.section_timeouts.get(name)} s; dumping stacks and terminating the rank\n")
                faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
                sys.stderr.flush()
                os._exit(int(os.environ.get("B200_FT_EXIT_CODE", "18")))

    def shutdown(self):
        self._stop.set()
        self._deadline = None

    # -- learned timeouts ---------------------------------------------------------------------
    def calculate_and_set_section_timeouts(self, selected_sections: List[str], calc_out_of_section: bool):
        names = list(selected_sections) + (["out_of_section"] if calc_out_of_section else [])
        if not names:
            return
        vals = torch.tensor([self.observed.get(n, 0.0) for n in names], dtype=torch.float64)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
            vals = vals.to(dev)
            dist.all_reduce(vals, op=dist.ReduceOp.MAX)
            vals = vals.cpu()
        with self._lock:
            for n, v in zip(names, vals.tolist()):
                if v > 0:
                    self.section_timeouts[n] = max(1.0, SAFETY_FACTOR * v)

    def state_dict(self) -> dict:
        return {"section_timeouts": dict(self.section_timeouts), "observed": dict(self.observed)}

    def load_state_dict(self, sd: dict):
        with self._lock:
            self.section_timeouts.update({k: float(v) for k, v in (sd.get("section_timeouts") or {}).items()})


_CLIENT: Optional[RankMonitor] = None
_S = {"state_path": None, "calc": False, "async": False, "setup_open": False, "iters": 0, "ckpts": 0, "persistent_loaded": False,
      "eval_idx": 0}


def _rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else int(os.environ.get("RANK", "0"))


def _print0(msg: str):
    if _rank() == 0:
        print(msg, flush=True)


def get_rank_monitor_client() -> Optional[RankMonitor]:
    return _CLIENT


def setup(args=None, timeouts: Optional[Dict[str, float]] = None) -> None:
    global _CLIENT
    save = getattr(args, "save", None)
    if args is not None and not save:
        raise ValueError("checkpointing save dir must be set to enable fault tolerance")
    if save and _rank() == 0:
        os.makedirs(save, exist_ok=True)                 # the timeouts file may be written before the first checkpoint
    if _CLIENT is not None:
        _CLIENT.shutdown()
    _CLIENT = RankMonitor(timeouts)
    _S.update(state_path=os.path.join(save, "ft_state.json") if save else None, calc=bool(getattr(args, "calc_ft_timeouts", False)),
              setup_open=True, iters=0, ckpts=0, persistent_loaded=False, eval_idx=0)
    _S["async"] = bool(getattr(args, "async_save", False))
    if _S["state_path"] and os.path.exists(_S["state_path"]):
        with open(_S["state_path"]) as f:
            _CLIENT.load_state_dict(json.load(f))
        _print0(f"FT: loaded timeouts from {_S['state_path']}: {_CLIENT.section_timeouts}")
    _print0(f"FT: initialized. Timeouts={_CLIENT.section_timeouts}")
    _CLIENT.start_section("setup")


def on_training_step_start() -> None:
    c = _CLIENT
    if c is None:
        return
    if _S["setup_open"]:
        c.end_section("setup")
        _S["setup_open"] = False
    if _S["iters"] >= NUM_WARMUP_ITERS:
        c.start_section("step")
    _S["eval_idx"] = 0


def on_training_step_end() -> None:
    c = _CLIENT
    if c is None:
        return
    if _S["iters"] >= NUM_WARMUP_ITERS:
        c.end_section("step")
    _S["iters"] += 1


def on_eval_step_start() -> None:
    c = _CLIENT
    if c is None:
        return
    if _S["setup_open"]:                                   # evaluation before any training (--skip-train)
        c.end_section("setup")
        _S["setup_open"] = False
    if _S["eval_idx"] >= NUM_WARMUP_ITERS:
        c.start_section("step")


def on_eval_step_end() -> None:
    c = _CLIENT
    if c is None:
        return
    if _S["eval_idx"] >= NUM_WARMUP_ITERS:
        c.end_section("step")
    _S["eval_idx"] += 1


def on_checkpointing_start() -> None:
    if _CLIENT is not None:
        _CLIENT.start_section("checkpointing")


def on_checkpointing_end(is_async_finalization: bool = False) -> None:
    c = _CLIENT
    if c is None:
        return
    c.end_section("checkpointing")
    # with async saves a checkpoint exists once it has been finalised, with sync saves right away
    if (not _S["async"]) or is_async_finalization:
        _S["ckpts"] += 1
        _maybe_update_timeouts()


def on_checkpoint_loaded(is_local_chkpt: bool = False) -> None:
    _S["persistent_loaded"] = not is_local_chkpt


def shutdown() -> None:
    global _CLIENT
    if _CLIENT is not None:
        _CLIENT.end_section()
        _maybe_update_timeouts(is_closing_ft=True)
        _CLIENT.shutdown()
        _CLIENT = None
        _print0("FT: closed.")


def _maybe_update_timeouts(is_closing_ft: bool = False) -> None:
    c = _CLIENT
    if c is None or not _S["calc"]:
        return
    sections = []
    if _S["persistent_loaded"]:
        sections.append("setup")
    if _S["iters"] - NUM_WARMUP_ITERS >= MIN_ITERS_FOR_STEP_TIMEOUT_UPDATE:
        sections.append("step")
    if _S["ckpts"] > 0 and not _S["async"]:
        sections.append("checkpointing")
    out_of_section = is_closing_ft and {"setup", "step"}.issubset(sections) and _S["ckpts"] > 0
    if not sections and not out_of_section:
        return
    c.calculate_and_set_section_timeouts(sections, out_of_section)
    if _rank() == 0 and _S["state_path"]:
        tmp = _S["state_path"] + ".tmp"
        with open(tmp, "w") as f:
            json.dump(c.state_dict(), f)
        os.replace(tmp, _S["state_path"])
    _print0(f"FT: updated timeouts for {sections}{' + out_of_section' if out_of_section else ''}: {c.section_timeouts}")


def maybe_setup_simulated_fault() -> None:
    """``FT_SIM_FAULT_DESC="rank_hung|rank_killed|random;<rank or empty>;<base delay s>"`` (testing only): after the delay (plus
    up to 50 % jitter) the chosen rank suspends itself with SIGSTOP (a hang: the watchdogs of its peers fire) or kills itself
    with SIGKILL (the launcher notices)."""
    desc = os.environ.get("FT_SIM_FAULT_DESC")
    if not desc:
        return
    kind, rank_to_fail, base_delay = (x.strip() for x in desc.split(";"))
    rng = random.Random()
    target = int(rank_to_fail) if rank_to_fail else rng.randint(0, (dist.get_world_size() if dist.is_initialized() else 1) - 1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor([target], device=dev)
        dist.broadcast(t, 0)
        target = int(t.item())
    _print0(f"FT: initializing simulated fault: {kind}, rank to fail: {target}, base delay: {base_delay}")
    if _rank() != target:
        return
    if kind == "random":
        kind = rng.choice(["rank_killed", "rank_hung"])
    sig = {"rank_killed": signal.SIGKILL, "rank_hung": signal.SIGSTOP}.get(kind)
    if sig is None:
        raise ValueError(f"unknown simulated fault type {kind!r}")
    delay = float(base_delay) * (1.0 + 0.5 * rng.random())

    def fire():
        time.sleep(delay)
        sys.stderr.write(f"\n####\nFT: simulating fault: {kind}; rank to fail: {target}\n####\n")
        sys.stderr.flush()
        os.kill(os.getpid(), sig)

    threading.Thread(target=fire, name="ft-simulated-fault", daemon=True).start()
