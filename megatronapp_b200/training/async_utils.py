"""Asynchronous checkpoint saves (reference ``training/async_utils.py`` over
``dist_checkpointing/strategies/async_utils.py``: ``init_persistent_async_worker``, ``schedule_async_save``,
``maybe_finalize_async_save``, ``is_empty_async_queue``).

Two kinds of work sit behind this front:

* ``torch_dist`` saves go to the long-lived *writer process* of ``ckpt/torch_dist.py`` (``AsyncWriter``: the trainer stages
  tensors in shared CPU memory and returns; the process writes the ``.distcp`` files and reports byte offsets) --
  ``init_persistent_async_worker`` starts it ahead of the first save (``--use-persistent-ckpt-worker``), so the first
  checkpoint does not pay the process spawn;
* legacy ``torch`` format files are written by per-file threads.

Finalisation (metadata, tracker file, "successfully saved") happens only after every rank's writes have completed --
``checkpointing.maybe_finalize_async_save`` agrees on that with an all-reduce."""
from . import checkpointing


def init_persistent_async_worker():
    """Start (or return) the writer process; idempotent."""
    from ..ckpt import torch_dist

    return torch_dist.get_async_writer()


def schedule_async_save(fn, *args, **kw):
    """Run ``fn(*args, **kw)`` (a file write) off the training thread; it is joined by ``maybe_finalize_async_save``."""
    import threading

    t = threading.Thread(target=fn, args=args, kwargs=kw, daemon=False)
    t.start()
    checkpointing._ASYNC_THREADS.append(t)
    return t


def maybe_finalize_async_save(blocking: bool = False, terminate: bool = False):
    checkpointing.maybe_finalize_async_save(blocking=blocking or terminate)
    if terminate:
        from ..ckpt import torch_dist

        torch_dist.shutdown_async_writer()


def is_empty_async_queue() -> bool:
    return not any(t.is_alive() for t in checkpointing._ASYNC_THREADS) and not checkpointing._PENDING
