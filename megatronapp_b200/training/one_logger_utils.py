"""End-to-end job metrics in the one-logger style (reference ``training/one_logger_utils.py``): app
start/finish, first-iteration latency, checkpoint save counts/durations, train/eval throughput --
collected in-process and written as JSON lines to ``--one-logger-dir`` (no external service)."""
import json
import os
import time
from typing import Dict

_M: Dict[str, float] = {}
_SINK = {"path": None}


def configure(path=None):
    _SINK["path"] = path
    _M.setdefault("app_start_time", time.time())


def _emit(event: str, **kw):
    rec = {"event": event, "time": time.time(), **kw}
    if _SINK["path"]:
        os.makedirs(os.path.dirname(_SINK["path"]) or ".", exist_ok=True)
        with open(_SINK["path"], "a") as f:
            f.write(json.dumps(rec) + "\n")
    return rec


def on_train_start(iteration, consumed_train_samples, train_samples, seq_length, train_iters, save, async_save,
                   log_throughput, num_floating_point_operations_so_far):
    _M.update(train_start_time=time.time(), start_iteration=iteration, seq_length=seq_length)
    return _emit("train_start", iteration=iteration, train_iters=train_iters)


def track_e2e_metrics(log_throughput=False, throughput=None):
    now = time.time()
    rec = {"e2e_seconds": now - _M.get("app_start_time", now)}
    if throughput is not None:
        rec["throughput"] = throughput
    return _emit("e2e", **rec)


def on_save_checkpoint_start(async_save):
    _M["save_start"] = time.time()
    _M["save_count"] = _M.get("save_count", 0) + 1


def on_save_checkpoint_end(save_checkpoint_duration=None, current_iteration=None, async_save=False):
    d = save_checkpoint_duration if save_checkpoint_duration is not None else time.time() - _M.get("save_start", time.time())
    _M["save_total"] = _M.get("save_total", 0.0) + d
    return _emit("save_checkpoint", seconds=d, iteration=current_iteration, count=_M.get("save_count", 0))


def on_pretrain_start():
    configure(_SINK["path"])
    return _emit("app_start")


def finish():
    return _emit("app_finish", total_seconds=time.time() - _M.get("app_start_time", time.time()))


def get_metrics():
    return dict(_M)
