"""Asynchronous checkpoint finalisation queue (reference ``training/async_utils.py``)."""
from . import checkpointing


def init_persistent_async_worker():
    return None          # writer threads are created lazily by checkpointing._write


def schedule_async_save(fn, *args, **kw):
    import threading

    t = threading.Thread(target=fn, args=args, kwargs=kw, daemon=False)
    t.start()
    checkpointing._ASYNC_THREADS.append(t)
    return t


def maybe_finalize_async_save(blocking: bool = False, terminate: bool = False):
    checkpointing.maybe_finalize_async_save(blocking=blocking or terminate)


def is_empty_async_queue() -> bool:
    return not any(t.is_alive() for t in checkpointing._ASYNC_THREADS)
