"""The reference's MegaDPP transport API (module functions of ``shm_tensor_new_rdma_pre_alloc`` /
``shm_tensor_new_rdma``: ``init_shared_memory``, ``put_forward_tensor(k, t)``, ``get_forward_tensor(k) -> Optional[Tensor]``,
``put_backward_tensor``, ``get_backward_tensor``, ``check_send_finish``, ``join_threads`` ...) on top of this framework's
transports, so schedulers written against the reference's extension keep working.

Reference semantics kept (``shm_tensor_new_rdma_pre_alloc.cpp:631,1476-1548``):

* slots are indexed ``k in [0, total_num_microbatches)`` (= microbatches x chunks, groups of ``pp`` microbatches per chunk);
* a tensor put by the LAST stage lands ``pp`` slots further on stage 0 (the next model chunk's slot); symmetrically a
  gradient put by the FIRST stage lands ``pp`` slots earlier on the last stage;
* ``get_*`` never blocks: it returns the tensor if it has arrived, else ``None`` -- that is what lets a scheduler poll
  for whichever input is ready;
* any dtype (the reference hard-codes fp16), no host staging: CUDA ranks use the NVLink mailbox
  (``csrc/mailbox.cpp``), CPU / gloo ranks tagged ``isend`` / ``irecv``.
"""
from __future__ import annotations

import threading
from typing import Dict, Optional

import torch
import torch.distributed as dist

_T: Optional["_Transport"] = None


class _Transport:
    def __init__(self, numel, rank, total, chunks, pp, prev_rank, next_rank, is_last, is_first, dtype, device):
        self.numel, self.rank, self.total, self.chunks, self.pp = numel, rank, total, chunks, pp
        self.prev, self.next, self.is_last, self.is_first = prev_rank, next_rank, is_last, is_first
        self.dtype, self.device = dtype, device
        self.sent = 0
        self.sends = []
        self.fwd: Dict[int, list] = {}
        self.bwd: Dict[int, list] = {}
        self.mailbox = None
        if device.type == "cuda" and dist.get_backend() == "nccl":
            from .scheduler import MailboxTransport

            self.mailbox = MailboxTransport(total, (numel,), dtype)
            self.mailbox.begin_iteration()
        else:
            for k in range(total):                      # a receive is always outstanding for every slot
                self._post(self.fwd, k, self.prev, k)
                self._post(self.bwd, k, self.next, total + k)

    def _post(self, table, k, src, tag):
        buf = torch.empty(self.numel, dtype=self.dtype, device=self.device)
        ev = threading.Event()
        work = dist.irecv(buf, src=src, tag=tag)

        def wait():
            try:
                work.wait()
                ev.set()
            except RuntimeError:                        # process group torn down while waiting
                pass

        threading.Thread(target=wait, daemon=True).start()
        table[k] = [ev, buf, src, tag]

    # -- reference API ---------------------------------------------------------------
    def put(self, k: int, t: torch.Tensor, forward: bool):
        if forward:
            slot, dst, tag0 = (k + self.pp if self.is_last else k) % self.total, self.next, 0
        else:
            slot, dst, tag0 = (k - self.pp if self.is_first else k) % self.total, self.prev, self.total
        flat = t.detach().reshape(-1).to(self.dtype)
        if self.mailbox is not None:
            (self.mailbox.put_forward if forward else self.mailbox.put_backward)(slot, flat)
        else:
            self.sends.append((dist.isend(flat.contiguous(), dst=dst, tag=tag0 + slot), flat))
        self.sent += 1

    def get(self, k: int, forward: bool) -> Optional[torch.Tensor]:
        if self.mailbox is not None:
            return (self.mailbox.try_get_forward if forward else self.mailbox.try_get_backward)(k)
        table = self.fwd if forward else self.bwd
        ev, buf, src, tag = table[k]
        if not ev.is_set():
            return None
        self._post(table, k, src, tag)                  # re-arm the slot for the next iteration
        return buf

    def close(self):
        """Drain the always-armed receives (each rank sends one dummy per slot to both neighbours) so that no thread is
        left blocked in ``wait()`` when the process group goes away."""
        if self.mailbox is not None:
            self.mailbox.shutdown()
            return
        self.all_sent()
        dist.barrier()
        z = torch.zeros(self.numel, dtype=self.dtype, device=self.device)
        hs = [dist.isend(z, dst=self.next, tag=k) for k in range(self.total)]
        hs += [dist.isend(z, dst=self.prev, tag=self.total + k) for k in range(self.total)]
        for h in hs:
            h.wait()
        for table in (self.fwd, self.bwd):
            for ev, *_ in table.values():
                ev.wait(timeout=30)

    def all_sent(self) -> bool:
        # every slot has a receive armed on the peer, so a pending send can always complete; gloo only reports
        # completion from inside wait(), hence the wait instead of a poll
        for h, _ in self.sends:
            h.wait()
        self.sends = []
        return True


def init_shared_memory(numel, rank, total_num_microbatches, num_model_chunks, pipeline_parallel_size, node_ips=None,
                       prev_rank=None, next_rank=None, is_pipeline_last_stage=False, is_pipeline_first_stage=False,
                       workload=None, num_gpus=None, dtype=torch.float16, device=None):
    """``node_ips`` / ``workload`` / ``num_gpus`` (sender-thread chunking, cross-node RDMA) are accepted and unused: one
    NVSwitch domain needs neither host staging nor chunked copies."""
    global _T
    device = torch.device(device) if device is not None else (
        torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
    _T = _Transport(int(numel), rank, int(total_num_microbatches), int(num_model_chunks), int(pipeline_parallel_size),
                    prev_rank, next_rank, bool(is_pipeline_last_stage), bool(is_pipeline_first_stage), dtype, device)


def init_forward_rdma(numel=None, rank=None, next_rank=None, prev_rank=None, pipeline_rank=None):
    """Cross-node queue pairs of the reference; nothing to set up inside one NVSwitch domain."""


init_backward_rdma = init_forward_rdma


def put_forward_tensor(k: int, t: torch.Tensor):
    _T.put(k, t, True)


def put_backward_tensor(k: int, t: torch.Tensor):
    _T.put(k, t, False)


def get_forward_tensor(k: int) -> Optional[torch.Tensor]:
    return _T.get(k, True)


def get_backward_tensor(k: int) -> Optional[torch.Tensor]:
    return _T.get(k, False)


def check_send_finish(forward_only: bool = False) -> bool:
    return _T.all_sent()


def thread_pool(*_a, **_k):
    """The older extension started its sender / receiver threads per iteration; receives are always armed here."""


def join_threads():
    global _T
    if _T is not None:
        _T.close()
    _T = None


clean_rdma = join_threads
