"""MegaDPP: dynamic pipeline planning.

The reference ships only the transport and flags for this feature; the scheduler
below is designed from the paper's description (``MegatronApp.pdf`` §5.2, ``README.md:101-111``)
and the committed traces (``profiling/*-dpp-aggregated.json``):

* the unit of work is a task ``F(m, c)`` / ``B(m, c)`` of the (microbatch x model-chunk) matrix;
* a task is *ready* when its input has arrived in the mailbox (``try_get_*`` is a non-blocking
  poll, exactly the primitive ``get_forward_tensor`` provides in the reference's C++ API);
* among ready tasks the policy picks
    ``bfc``  breadth-first: same chunk, next microbatch  (earliest grad-sync, least p2p pressure;
             this is the order the committed DPP traces show on stage 0),
    ``dfc``  depth-first: drain backwards first, then the same microbatch's next chunk
             (1F1B-like, lowest activation memory),
    ``auto`` BFC until the number of stashed activations reaches the static interleaved
             schedule's peak, then DFC ("BFC best-effort until OOM risk"),
    ``static`` the upstream interleaved order over the mailbox transport (A/B baseline);
* invariants (unit-tested with a single-process simulator): every (m, c) runs F then B exactly
  once, the stash is bounded under ``dfc``/``auto``, a chunk's grad-sync is released by its
  last backward, and no order can deadlock because sends never block and receives are polls.

Slot index of a task is ``c * M + m`` on every rank, so neighbours need not agree on an order.
"""
from __future__ import annotations

import contextlib
import time
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from ..parallel import schedules
from ..parallel import state as ps
from ..parallel.schedules import Task, backward_step, forward_step, get_tensor_shapes, interleaved_warmup
from ..scan.tracer import get_tracer

_TRANSPORT = None
_POLICY = "auto"


# ---------------------------------------------------------------------------
# policy (pure, unit-testable)
# ---------------------------------------------------------------------------

def pick_task(ready_f: List[Tuple[int, int]], ready_b: List[Tuple[int, int]], policy: str, stash: int,
              stash_limit: int, head: Optional[Tuple[int, int]] = None) -> Optional[Task]:
    """``ready_*`` are lists of (microbatch, chunk).  Returns the next task or None.

    ``auto`` is the memory-safe policy: forwards are admitted strictly in the static interleaved
    table order (``head`` = first not-yet-run entry of ``get_schedule_table``) and only while the
    stash is below the static schedule's in-flight bound -- that bound is only meaningful with
    head-of-line forward order (admitting *any* ready forward can fill the stash with later-group
    chunk-0 work and starve the deeper-chunk forwards every backward depends on).  Backwards are
    fully dynamic: any ready backward runs as soon as the head forward is blocked or the bound is
    reached, which is where the ready-queue wins over a fixed order when a stage is slow."""
    if not ready_f and not ready_b:
        return None
    if policy == "auto":
        head_ready = head is not None and head in ready_f and stash < stash_limit
        if ready_b and not head_ready:
            m, c = min(ready_b, key=lambda t: (t[0], -t[1]))
            return Task("B", m, c)
        if head_ready:
            if ready_b and stash >= stash_limit - 1:
                m, c = min(ready_b, key=lambda t: (t[0], -t[1]))
                return Task("B", m, c)
            return Task("F", head[0], head[1])
        return None
    if policy == "bfc":
        if ready_f:
            m, c = min(ready_f, key=lambda t: (t[1], t[0]))      # lowest chunk, then microbatch
            return Task("F", m, c)
        m, c = min(ready_b, key=lambda t: (-t[1], t[0]))         # deepest chunk first on the way back
        return Task("B", m, c)
    # dfc
    if ready_b:
        m, c = min(ready_b, key=lambda t: (t[0], -t[1]))
        return Task("B", m, c)
    m, c = min(ready_f, key=lambda t: (t[0], t[1]))
    return Task("F", m, c)


def _head_of_line(table, f_done):
    for mc in table:
        if mc not in f_done:
            return mc
    return None


def simulate(pp: int, M: int, C: int, policy: str, latency: int = 1, stash_limit: Optional[int] = None,
             f_cost: int = 1, b_cost: int = 2, slow_rank: Optional[int] = None, slow_factor: int = 1):
    """Discrete-time single-process simulation of ``pp`` ranks running the dynamic scheduler.
    Returns (per-rank task order, makespan, per-rank peak stash).  Used by the tests to prove
    completeness / boundedness / deadlock-freedom without a process group."""
    fwd_arrival: List[Dict[Tuple[int, int], int]] = [dict() for _ in range(pp)]
    bwd_arrival: List[Dict[Tuple[int, int], int]] = [dict() for _ in range(pp)]
    f_done = [set() for _ in range(pp)]
    b_done = [set() for _ in range(pp)]
    busy_until = [0] * pp
    order: List[List[Task]] = [[] for _ in range(pp)]
    peak = [0] * pp
    total = M * C
    t = 0
    table = schedules.get_schedule_table(M, C, min(pp, M))
    limits = [stash_limit if stash_limit is not None else interleaved_warmup(pp, r, M, C, min(pp, M)) + 1
              for r in range(pp)]
    guard = 0
    while any(len(b_done[r]) < total for r in range(pp)):
        progressed = False
        for r in range(pp):
            if busy_until[r] > t or len(b_done[r]) == total:
                continue
            rf, rb = [], []
            for c in range(C):
                for m in range(M):
                    first = (r == 0 and c == 0)
                    last = (r == pp - 1 and c == C - 1)
                    if (m, c) not in f_done[r]:
                        if first or fwd_arrival[r].get((m, c), 10 ** 9) <= t:
                            rf.append((m, c))
                    elif (m, c) not in b_done[r]:
                        if last or bwd_arrival[r].get((m, c), 10 ** 9) <= t:
                            rb.append((m, c))
            task = pick_task(rf, rb, policy, len(f_done[r]) - len(b_done[r]), limits[r],
                             _head_of_line(table, f_done[r]))
            if task is None:
                continue
            progressed = True
            cost = (f_cost if task.kind == "F" else b_cost) * (slow_factor if r == slow_rank else 1)
            busy_until[r] = t + cost
            order[r].append(task)
            m, c = task.microbatch, task.chunk
            if task.kind == "F":
                f_done[r].add((m, c))
                peak[r] = max(peak[r], len(f_done[r]) - len(b_done[r]))
                if not (r == pp - 1 and c == C - 1):
                    nr, nc = (r + 1, c) if r < pp - 1 else (0, c + 1)
                    fwd_arrival[nr][(m, nc)] = t + cost + latency
            else:
                b_done[r].add((m, c))
                if not (r == 0 and c == 0):
                    pr, pc = (r - 1, c) if r > 0 else (pp - 1, c - 1)
                    bwd_arrival[pr][(m, pc)] = t + cost + latency
        t += 1
        guard += 1
        if guard > 100000:
            raise RuntimeError("simulation did not terminate (deadlock)")
    return order, max(busy_until), peak


# ---------------------------------------------------------------------------
# transports
# ---------------------------------------------------------------------------

class MailboxTransport:
    """NVLink peer mailbox (native ``_C.mailbox_*``)."""

    def __init__(self, n_slots: int, shape, dtype):
        from .. import ops

        self.ext = ops.ext()
        self.shape, self.dtype = tuple(shape), dtype
        nbytes = torch.empty(0, dtype=dtype).element_size()
        for s in shape:
            nbytes *= s
        self.epoch = 0
        me = dist.get_rank()
        self.dev = torch.cuda.current_device()
        job = f"{dist.get_world_size()}_{abs(hash((os_getpid_group(),))) % 10 ** 8}"
        self.fwd_id, hf = self.ext.mailbox_create(n_slots, nbytes, f"/b200dpp_{job}_f{me}")
        self.bwd_id, hb = self.ext.mailbox_create(n_slots, nbytes, f"/b200dpp_{job}_b{me}")
        info = [None] * dist.get_world_size()
        dist.all_gather_object(info, (me, hf, f"/b200dpp_{job}_f{me}", hb, f"/b200dpp_{job}_b{me}"))
        by_rank = {x[0]: x for x in info}
        nxt, prv = ps.get_pipeline_model_parallel_next_rank(), ps.get_pipeline_model_parallel_prev_rank()
        # my fwd box *sends* into next rank's fwd box; my bwd box sends into prev rank's bwd box
        self.ext.mailbox_connect(self.fwd_id, by_rank[nxt][1], by_rank[nxt][2])
        self.ext.mailbox_connect(self.bwd_id, by_rank[prv][3], by_rank[prv][4])
        dist.barrier()

    def begin_iteration(self, expected_fwd=None, expected_bwd=None):
        self.epoch += 1

    def put_forward(self, slot, t):
        self.ext.mailbox_put(self.fwd_id, slot, t.contiguous(), self.epoch)

    def put_backward(self, slot, t):
        self.ext.mailbox_put(self.bwd_id, slot, t.contiguous(), self.epoch)

    def try_get_forward(self, slot):
        if self.ext.mailbox_try_get(self.fwd_id, slot, self.epoch):
            return self.ext.mailbox_slot_tensor(self.fwd_id, slot, list(self.shape), self.dtype, self.dev)
        return None

    def try_get_backward(self, slot):
        if self.ext.mailbox_try_get(self.bwd_id, slot, self.epoch):
            return self.ext.mailbox_slot_tensor(self.bwd_id, slot, list(self.shape), self.dtype, self.dev)
        return None

    def all_sent(self) -> bool:
        return True  # puts are stream-ordered; completion == stream completion

    def end_iteration(self):
        torch.cuda.synchronize()
        dist.barrier(group=ps.get_pipeline_model_parallel_group())

    def shutdown(self):
        self.ext.mailbox_destroy(self.fwd_id)
        self.ext.mailbox_destroy(self.bwd_id)


def os_getpid_group():
    import os

    return os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "0")


class TaggedP2PTransport:
    """Same API over torch.distributed tagged isend/irecv (gloo / CPU, and the NCCL fallback when
    peers are on different nodes): receives for every expected slot are pre-posted at iteration
    start and polled with ``is_completed``."""

    def __init__(self, n_slots: int, shape, dtype, device):
        self.shape, self.dtype, self.device = tuple(shape), dtype, device
        self.n_slots = n_slots
        self.fwd_req, self.bwd_req, self.sends = {}, {}, []

    def begin_iteration(self, expected_fwd, expected_bwd):
        prv, nxt = ps.get_pipeline_model_parallel_prev_rank(), ps.get_pipeline_model_parallel_next_rank()
        self.fwd_req, self.bwd_req, self.sends = {}, {}, []
        for slot in expected_fwd:
            buf = torch.empty(self.shape, dtype=self.dtype, device=self.device)
            self.fwd_req[slot] = (self._watch(dist.irecv(buf, src=prv, tag=slot)), buf)
        for slot in expected_bwd:
            buf = torch.empty(self.shape, dtype=self.dtype, device=self.device)
            self.bwd_req[slot] = (self._watch(dist.irecv(buf, src=nxt, tag=self.n_slots + slot)), buf)

    @staticmethod
    def _watch(work):
        """gloo only completes a receive inside ``wait()``: a daemon thread per outstanding
        request turns that into a pollable flag (this transport is the CPU / cross-node
        fallback; the NVLink mailbox polls a host-visible doorbell word instead)."""
        import threading

        ev = threading.Event()

        def run():
            work.wait()
            ev.set()

        threading.Thread(target=run, daemon=True).start()
        return ev

    def put_forward(self, slot, t):
        self.sends.append((dist.isend(t.contiguous(), dst=ps.get_pipeline_model_parallel_next_rank(), tag=slot), t))

    def put_backward(self, slot, t):
        self.sends.append((dist.isend(t.contiguous(), dst=ps.get_pipeline_model_parallel_prev_rank(),
                                      tag=self.n_slots + slot), t))

    def _poll(self, table, slot):
        r = table.get(slot)
        if r is not None and r[0].is_set():
            return r[1]
        return None

    def try_get_forward(self, slot):
        return self._poll(self.fwd_req, slot)

    def try_get_backward(self, slot):
        return self._poll(self.bwd_req, slot)

    def all_sent(self):
        return all(h.is_completed() for h, _ in self.sends)

    def end_iteration(self):
        for h, _ in self.sends:
            h.wait()
        self.sends = []

    def shutdown(self):
        pass


# ---------------------------------------------------------------------------
# the dynamic executor
# ---------------------------------------------------------------------------

class _CachedIterator:
    """Random access by microbatch index over a sequential data iterator."""

    def __init__(self, it):
        self.it, self.cache = it, []

    def at(self, m):
        while len(self.cache) <= m:
            self.cache.append(next(self.it))
        return iter([self.cache[m]])


def dpp_forward_backward(*, forward_step_func, data_iterator, model, num_microbatches, seq_length,
                         micro_batch_size, decoder_seq_length=None, forward_only=False,
                         collect_non_loss_data=False, first_val_step=None):
    chunks = model if isinstance(model, (list, tuple)) else [model]
    C, M = len(chunks), num_microbatches
    pp, rank = ps.get_pipeline_model_parallel_world_size(), ps.get_pipeline_model_parallel_rank()
    config = schedules._get_config(chunks[0])
    if pp == 1:
        return schedules.forward_backward_no_pipelining(
            forward_step_func=forward_step_func, data_iterator=data_iterator, model=model,
            num_microbatches=M, seq_length=seq_length, micro_batch_size=micro_batch_size,
            forward_only=forward_only, collect_non_loss_data=collect_non_loss_data)
    tr = _TRANSPORT
    iters = data_iterator if isinstance(data_iterator, (list, tuple)) else [data_iterator] * C
    cached = [(_CachedIterator(it) if it is not None else None) for it in iters]
    slot = lambda m, c: c * M + m  # noqa: E731
    first = lambda c: rank == 0 and c == 0  # noqa: E731
    last = lambda c: rank == pp - 1 and c == C - 1  # noqa: E731
    exp_f = [slot(m, c) for c in range(C) for m in range(M) if not first(c)]
    exp_b = [] if forward_only else [slot(m, c) for c in range(C) for m in range(M) if not last(c)]
    tr.begin_iteration(exp_f, exp_b)
    f_done, b_done, stash = set(), set(), {}
    store: List = []
    total_tokens = torch.zeros((), dtype=torch.int, device=next(chunks[0].parameters()).device)
    # in-flight cap (forwards not yet matched by a backward); evaluation has no backwards to drain it, so no cap
    limit = (M * C + 1) if forward_only else interleaved_warmup(pp, rank, M, C, min(pp, M)) + 1
    table = schedules.get_schedule_table(M, C, min(pp, M))
    b_count = [0] * C
    tracer = get_tracer()
    static_order = None
    if _POLICY == "static":
        static_order = (schedules.schedule_interleaved(pp, rank, M, C, min(pp, M), forward_only) if C > 1
                        else schedules.schedule_1f1b(pp, rank, M, forward_only))
    for ch in chunks:
        if hasattr(ch, "is_last_microbatch"):
            ch.is_last_microbatch = False
    n_total = M * C
    spins = 0
    while len(f_done) < n_total or (not forward_only and len(b_done) < n_total):
        arrived_f, arrived_b = {}, {}
        if static_order is not None:
            t = static_order[len(f_done) + len(b_done)]
            while True:  # blocking wait on exactly this task's input
                if t.kind == "F":
                    x = None if first(t.chunk) else tr.try_get_forward(slot(t.microbatch, t.chunk))
                    if first(t.chunk) or x is not None:
                        arrived_f[(t.microbatch, t.chunk)] = x
                        break
                else:
                    g = None if last(t.chunk) else tr.try_get_backward(slot(t.microbatch, t.chunk))
                    if last(t.chunk) or g is not None:
                        arrived_b[(t.microbatch, t.chunk)] = g
                        break
                time.sleep(0)
            task = t
        else:
            rf, rb = [], []
            for c in range(C):
                for m in range(M):
                    if (m, c) not in f_done:
                        x = None if first(c) else tr.try_get_forward(slot(m, c))
                        if first(c) or x is not None:
                            rf.append((m, c)); arrived_f[(m, c)] = x
                    elif not forward_only and (m, c) not in b_done:
                        g = None if last(c) else tr.try_get_backward(slot(m, c))
                        if last(c) or g is not None:
                            rb.append((m, c)); arrived_b[(m, c)] = g
            task = pick_task(rf, rb, _POLICY, len(f_done) - len(b_done), limit, _head_of_line(table, f_done))
            if task is None:
                spins += 1
                time.sleep(0 if spins < 1000 else 0.0002)
                continue
        spins = 0
        m, c = task.microbatch, task.chunk
        if C > 1:
            ps.set_virtual_pipeline_model_parallel_rank(c)
        if task.kind == "F":
            inp = arrived_f[(m, c)]
            if inp is not None:
                inp = inp.detach().clone().requires_grad_(not forward_only)
            it = cached[c].at(m) if cached[c] is not None else None
            ctx = torch.no_grad() if forward_only else contextlib.nullcontext()
            with ctx:
                out, ntok = forward_step(forward_step_func, it, chunks[c], M, inp, store, config,
                                         collect_non_loss_data, is_first_microbatch=(m == 0),
                                         current_microbatch=m, is_last_stage=last(c))
            total_tokens += ntok.to(total_tokens.device) if torch.is_tensor(ntok) else int(ntok)
            if not last(c):
                nc = c if rank < pp - 1 else c + 1
                tr.put_forward(slot(m, nc), out.detach())
            if not forward_only:
                stash[(m, c)] = (inp, out)
            f_done.add((m, c))
        else:
            inp, out = stash.pop((m, c))
            grad = arrived_b[(m, c)]
            if grad is not None:
                grad = grad.detach().clone()
            b_count[c] += 1
            if hasattr(chunks[c], "is_last_microbatch"):
                chunks[c].is_last_microbatch = (b_count[c] == M)
            if b_count[c] == M:
                tracer.instant("reduce", model_chunk=c)
            in_grad = backward_step(inp, out, grad, None, config, microbatch=m)
            if not first(c):
                pc = c if rank > 0 else c - 1
                tr.put_backward(slot(m, pc), in_grad)
            b_done.add((m, c))
    tr.end_iteration()
    if C > 1:
        ps.set_virtual_pipeline_model_parallel_rank(0)
    for ch in chunks:
        if hasattr(ch, "is_last_microbatch"):
            ch.is_last_microbatch = True
    if not forward_only and config.finalize_model_grads_func is not None:
        config.finalize_model_grads_func(chunks, total_tokens if config.calculate_per_token_loss else None)
    return store


def make_transport(seq_length, micro_batch_size, config, num_microbatches, num_chunks, device=None):
    shape = get_tensor_shapes(seq_length, micro_batch_size, config)
    dtype = config.pipeline_dtype or config.params_dtype
    n_slots = num_microbatches * num_chunks
    if torch.cuda.is_available() and dist.get_backend() == "nccl":
        return MailboxTransport(n_slots, shape, dtype)
    return TaggedP2PTransport(n_slots, shape, dtype, device or "cpu")


def install_dpp_schedule(args=None, policy: Optional[str] = None, transport=None):
    """``--use-dpp``: route ``get_forward_backward_func()`` to the dynamic scheduler."""
    global _TRANSPORT, _POLICY
    _POLICY = policy or (getattr(args, "dpp_policy", None) or "auto")
    _TRANSPORT = transport
    schedules.set_schedule_override(_lazy_dpp)


def _lazy_dpp(**kw):
    global _TRANSPORT
    if _TRANSPORT is None and ps.get_pipeline_model_parallel_world_size() > 1:
        chunks = kw["model"] if isinstance(kw["model"], (list, tuple)) else [kw["model"]]
        cfg = schedules._get_config(chunks[0])
        _TRANSPORT = make_transport(kw["seq_length"], kw["micro_batch_size"], cfg, kw["num_microbatches"],
                                    len(chunks), next(chunks[0].parameters()).device)
    return dpp_forward_backward(**kw)


def shutdown_dpp():
    global _TRANSPORT
    if _TRANSPORT is not None:
        _TRANSPORT.shutdown()
        _TRANSPORT = None
    schedules.set_schedule_override(None)
