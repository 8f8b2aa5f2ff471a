"""MegaFBD: forward / backward disaggregation onto separate devices.

Capability parity with reference ``schedules.py:2208-2506``
(``forward_or_backward_pipelining_without_interleaving``), ``parallel_state.py:444-512`` (F/B
pairing inside DP) and ``p2p_communication.py:723-756`` (stage-input hand-off), re-designed:

* each DP pair = (forward rank, backward rank) holding the same pipeline stage / TP shard;
* forward ranks run a forward-only pipeline under ``no_grad`` and ship every stage's *input*
  activation to their backward twin over a dedicated unidirectional channel (the last stage
  only hands its input over -- its own forward would be thrown away);
* backward ranks receive the stage input, recompute the stage forward with grad, receive the
  output gradient from the next backward stage, run backward and pass the input gradient on:
  they hold at most one microbatch of activations;
* gradients are averaged over the backward ranks only (``half_dp`` group, so
  ``get_data_parallel_world_size()`` reports dp/2 exactly like the reference :2851-2864);
* instead of forward ranks carrying zero-gradient all-reduces and a redundant optimizer
  (SURVEY §7.4-7), the backward rank steps the optimizer and broadcasts the updated parameter
  buffer to its forward twin over NVLink (2 bytes/param/step) -- ``sync_forward_weights``;
* there is no coordinator process and no thread-level "virtual TP" shim (``Controller.py``,
  ``virtual_tensor_parallel_communication.py``): the channel transport of ``parallel/p2p.py``
  is deadlock-free for this data flow by construction.

``--ignore-forward-tensor-parallel`` (reference ``parallel_state.py:896-964``): every (stage, replica) becomes a
block of ``1 + tp`` ranks -- ONE forward rank that holds the un-sharded stage and runs true TP=1 math (the reference
emulates the tp virtual ranks with threads on that GPU and a gloo "virtual TP" shim; here it is simply a TP=1 model),
plus the ``tp`` backward ranks.  The forward rank feeds each backward rank its (sequence-parallel slice of the) stage
input and, after every optimizer step, re-assembles its weights from the ``tp`` updated shards
(``sync_forward_weights``: column-parallel shards concatenate along dim 0, row-parallel along dim 1, gated fc1
re-interleaves ``[gate_r; up_r]`` into ``[gate; up]``, replicated parameters come from TP rank 0).
"""
from __future__ import annotations

import contextlib
from typing import List

import torch
import torch.distributed as dist

from ..parallel import p2p, schedules
from ..parallel import state as ps
from ..parallel.schedules import backward_step, forward_step, get_tensor_shapes


def fbd_forward_backward(*, forward_step_func, data_iterator, model, num_microbatches, seq_length,
                         micro_batch_size, decoder_seq_length=None, forward_only=False,
                         collect_non_loss_data=False, first_val_step=None):
    chunks = model if isinstance(model, (list, tuple)) else [model]
    if len(chunks) != 1:
        return _fbd_interleaved(forward_step_func, data_iterator, chunks, num_microbatches, seq_length, micro_batch_size,
                                forward_only, collect_non_loss_data)
    m0 = chunks[0]
    config = schedules._get_config(m0)
    pp, rank = ps.get_pipeline_model_parallel_world_size(), ps.get_pipeline_model_parallel_rank()
    shape = get_tensor_shapes(seq_length, micro_batch_size, config)
    bwd_config = config
    if ps.fbd_forward_tp_collapsed() and ps.is_forward_stage():
        # this rank runs TP=1 / no SP; what it ships must have the layout its BACKWARD twins expect
        bwd_config = getattr(config, "fbd_backward_config", config)
    dtype = config.pipeline_dtype or config.params_dtype
    first, last = rank == 0, rank == pp - 1
    it = data_iterator[0] if isinstance(data_iterator, (list, tuple)) else data_iterator
    store: List = []
    total_tokens = torch.zeros((), dtype=torch.int, device=next(m0.parameters()).device)
    if ps.is_forward_stage():
        # ---------------- forward-only pipeline, hand stage inputs to the twin
        for mb in range(num_microbatches):
            inp = None
            if not first:
                inp = p2p.recv(shape, dtype, ps.get_pipeline_model_parallel_prev_rank(), "fwd", "recv-forward")
                p2p.send_corresponding_forward(inp, bwd_config)
            if last:
                if it is not None and pp > 1:
                    pass  # labels are read by the backward twin from its own (identical) iterator
                continue
            with torch.no_grad():
                out, _ = forward_step(forward_step_func, it, m0, num_microbatches, inp, store, config,
                                      collect_non_loss_data, is_first_microbatch=(mb == 0),
                                      current_microbatch=mb, is_last_stage=False)
            p2p.send(out, ps.get_pipeline_model_parallel_next_rank(), "fwd", "send-forward")
        p2p.flush()
        return []
    # ---------------- backward ranks: recompute + backward, one microbatch of activations at a time
    if hasattr(m0, "is_last_microbatch"):
        m0.is_last_microbatch = False
    for mb in range(num_microbatches):
        inp = None
        if not first:
            inp = p2p.recv_corresponding_forward(shape, dtype)
        ctx = torch.no_grad() if forward_only else contextlib.nullcontext()
        with ctx:
            out, ntok = forward_step(forward_step_func, it, m0, num_microbatches, inp, store, config,
                                     collect_non_loss_data, is_first_microbatch=(mb == 0),
                                     current_microbatch=mb, is_last_stage=last)
        total_tokens += ntok.to(total_tokens.device) if torch.is_tensor(ntok) else int(ntok)
        if forward_only:
            continue
        grad = None
        if not last:
            grad = p2p.recv(shape, dtype, ps.get_pipeline_model_parallel_next_rank(), "bwd", "recv-backward")
        if hasattr(m0, "is_last_microbatch"):
            m0.is_last_microbatch = (mb == num_microbatches - 1)
        in_grad = backward_step(inp, out, grad, None, config, microbatch=mb)
        if not first:
            p2p.send(in_grad, ps.get_pipeline_model_parallel_prev_rank(), "bwd", "send-backward")
    p2p.flush()
    if hasattr(m0, "is_last_microbatch"):
        m0.is_last_microbatch = True
    if not forward_only and config.finalize_model_grads_func is not None:
        config.finalize_model_grads_func(chunks, total_tokens if config.calculate_per_token_loss else None)
    return store


def _fbd_interleaved(forward_step_func, data_iterator, chunks, num_microbatches, seq_length, micro_batch_size,
                     forward_only, collect_non_loss_data):
    """MegaFBD on the interleaved (virtual pipeline) layout: rank r holds chunks c = 0..V-1, i.e. virtual stages
    ``c * pp + r``.  Microbatches travel in groups of ``pp`` (the interleaved unit):

    * forward ranks run group g through chunk 0, then chunk 1, ... (the order of the interleaved warm-up, so the ring
      stays busy) and hand every virtual stage's input to their twin in that order;
    * a backward rank first collects the group's ``V x pp`` stage inputs from its twin (same FIFO order; ~one activation
      tensor each), then walks the virtual stages of the group *backwards* -- chunk V-1 ... 0 -- recomputing one
      microbatch at a time, so it still holds a single microbatch of activations.

    Every directed edge carries its messages in one agreed order ((group, chunk ascending, microbatch) for activations,
    (group, chunk descending, microbatch) for gradients, the wrap edge last->first included), so the unidirectional
    channels cannot dead-lock."""
    V = len(chunks)
    pp, rank = ps.get_pipeline_model_parallel_world_size(), ps.get_pipeline_model_parallel_rank()
    if num_microbatches % pp:
        raise RuntimeError(f"interleaved MegaFBD needs num_microbatches ({num_microbatches}) divisible by pp ({pp})")
    config = schedules._get_config(chunks[0])
    shape = get_tensor_shapes(seq_length, micro_batch_size, config)
    bwd_config = config
    if ps.fbd_forward_tp_collapsed() and ps.is_forward_stage():
        bwd_config = getattr(config, "fbd_backward_config", config)
    dtype = config.pipeline_dtype or config.params_dtype
    its = list(data_iterator) if isinstance(data_iterator, (list, tuple)) else [data_iterator] * V
    nxt, prv = ps.get_pipeline_model_parallel_next_rank(), ps.get_pipeline_model_parallel_prev_rank()
    is_first = lambda c: rank == 0 and c == 0                       # noqa: E731  first / last VIRTUAL stage
    is_last = lambda c: rank == pp - 1 and c == V - 1               # noqa: E731
    groups = [list(range(g, g + pp)) for g in range(0, num_microbatches, pp)]
    store: List = []
    total_tokens = torch.zeros((), dtype=torch.int, device=next(chunks[0].parameters()).device)
    if ps.is_forward_stage():
        for grp in groups:
            for c in range(V):
                ps.set_virtual_pipeline_model_parallel_rank(c)
                for mb in grp:
                    inp = None
                    if not is_first(c):
                        inp = p2p.recv(shape, dtype, prv, "fwd", "recv-forward")
                        p2p.send_corresponding_forward(inp, bwd_config)
                    if is_last(c):
                        continue
                    with torch.no_grad():
                        out, _ = forward_step(forward_step_func, its[c], chunks[c], num_microbatches, inp, store, config,
                                              collect_non_loss_data, is_first_microbatch=(mb == 0), current_microbatch=mb,
                                              is_last_stage=False)
                    p2p.send(out, nxt, "fwd", "send-forward")
        p2p.flush()
        ps.set_virtual_pipeline_model_parallel_rank(0)
        return []
    for ch in chunks:
        if hasattr(ch, "is_last_microbatch"):
            ch.is_last_microbatch = False
    for grp in groups:
        stash = {}
        for c in range(V):
            for mb in grp:
                if not is_first(c):
                    stash[(c, mb)] = p2p.recv_corresponding_forward(shape, dtype)
        for c in reversed(range(V)):
            ps.set_virtual_pipeline_model_parallel_rank(c)
            for mb in grp:
                inp = stash.pop((c, mb), None)
                ctx = torch.no_grad() if forward_only else contextlib.nullcontext()
                with ctx:
                    out, ntok = forward_step(forward_step_func, its[c], chunks[c], num_microbatches, inp, store, config,
                                             collect_non_loss_data, is_first_microbatch=(mb == 0), current_microbatch=mb,
                                             is_last_stage=is_last(c))
                total_tokens += ntok.to(total_tokens.device) if torch.is_tensor(ntok) else int(ntok)
                if forward_only:
                    continue
                grad = None if is_last(c) else p2p.recv(shape, dtype, nxt, "bwd", "recv-backward")
                if hasattr(chunks[c], "is_last_microbatch"):
                    chunks[c].is_last_microbatch = (mb == num_microbatches - 1)
                in_grad = backward_step(inp, out, grad, None, config, microbatch=mb)
                if not is_first(c):
                    p2p.send(in_grad, prv, "bwd", "send-backward")
    p2p.flush()
    ps.set_virtual_pipeline_model_parallel_rank(0)
    for ch in chunks:
        if hasattr(ch, "is_last_microbatch"):
            ch.is_last_microbatch = True
    if not forward_only and config.finalize_model_grads_func is not None:
        config.finalize_model_grads_func(chunks, total_tokens if config.calculate_per_token_loss else None)
    return store


def _is_gated_fc1(name: str, config) -> bool:
    return bool(getattr(config, "gated_linear_unit", False)) and name.endswith(("mlp.linear_fc1.weight", "mlp.linear_fc1.bias"))


def _sync_forward_weights_collapsed(model) -> None:
    """tp backward shards -> one un-sharded forward copy.  Both sides walk ``named_parameters()`` in the same order;
    every backward rank ships ONE flat buffer (all its parameters back to back) over its ``fbw`` channel."""
    st = ps.get_state()
    block = st.ranks("fb")
    fwd, twins = block[0], block[1:]
    tp = len(twins)
    chunks = model if isinstance(model, (list, tuple)) else [model]
    for ch in chunks:
        inner = ch
        while hasattr(inner, "module"):
            inner = inner.module
        params = list(inner.named_parameters())
        if ps.is_backward_stage():
            flat = torch.cat([p.detach().reshape(-1) for _, p in params])
            p2p.send(flat, fwd, "fbw", "fbd-weights-send")
            continue
        config = inner.config
        dtype = params[0][1].dtype
        # shard sizes as the backward ranks see them
        sizes = []
        for name, p in params:
            sharded = getattr(p, "tensor_model_parallel", False)
            sizes.append(p.numel() // tp if sharded else p.numel())
        total = sum(sizes)
        dev = params[0][1].device
        shards = [p2p.recv((total,), dtype, src, "fbw", "fbd-weights-recv", device=dev) for src in twins]
        off = 0
        with torch.no_grad():
            for (name, p), n in zip(params, sizes):
                if not getattr(p, "tensor_model_parallel", False):
                    p.copy_(shards[0][off:off + n].view_as(p))
                else:
                    dim = getattr(p, "partition_dim", 0)
                    shp = list(p.shape)
                    shp[dim] //= tp
                    parts = [s_[off:off + n].view(shp) for s_ in shards]
                    if _is_gated_fc1(name, config):
                        half = shp[0] // 2       # every shard is [gate_r; up_r]: the full tensor is [gate_0..; up_0..]
                        full = torch.cat([x[:half] for x in parts] + [x[half:] for x in parts], dim=0)
                    else:
                        full = torch.cat(parts, dim=dim)
                    p.copy_(full)
                off += n
    p2p.flush()


def sync_forward_weights(model) -> None:
    """Backward rank -> forward twin broadcast of the (updated) contiguous parameter buffers."""
    st = ps.get_state()
    if not st.topo.fbd:
        return
    if st.topo.fbd_collapse:
        _sync_forward_weights_collapsed(model)
        return
    g = st.group("fb")
    pair = st.ranks("fb")
    src = pair[1]  # odd DP index = backward rank = owner of the optimizer
    chunks = model if isinstance(model, (list, tuple)) else [model]
    for ch in chunks:
        bufs = getattr(ch, "buffers", None)
        if bufs is None:
            for p in ch.parameters():
                dist.broadcast(p.data, src=src, group=g)
        else:
            for b in list(ch.buffers) + list(getattr(ch, "expert_parallel_buffers", [])):
                dist.broadcast(b.param_data, src=src, group=g)


def install_fbd_schedule(args=None) -> None:
    schedules.set_schedule_override(fbd_forward_backward)
