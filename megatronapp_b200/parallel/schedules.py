"""Forward/backward schedules: no pipelining, 1F1B, interleaved 1F1B -- as *task lists*.

API parity with reference ``core/pipeline_parallel/schedules.py``:
``get_forward_backward_func()`` (:28-117) returns a callable taking ``forward_step_func,
data_iterator, model, num_microbatches, seq_length, micro_batch_size, decoder_seq_length,
forward_only, collect_non_loss_data, first_val_step``; ``forward_step`` / ``backward_step``
(:179-303, :533-610) keep the ``(output_tensor, loss_func)`` protocol; warm-up counts and the
interleaved visiting order follow ``get_schedule_table`` (:758-853).

Design: a schedule is a pure function ``(pp, rank, M, C, G) -> [Task]`` (unit-testable without a
process group); ONE executor runs any task list over the unidirectional p2p channels of
``parallel/p2p.py``.  MegaDPP (``dpp/scheduler.py``) swaps the static list for a ready-queue,
MegaFBD (``fbd/schedule.py``) splits it across forward/backward twins -- same executor.
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass
from typing import Callable, Dict, Iterator, List, Optional, Tuple, Union

import torch

from . import p2p
from . import state as ps
from ..scan.tracer import scope as _scope


@dataclass(frozen=True)
class Task:
    kind: str        # "F" | "B"
    microbatch: int
    chunk: int = 0


# ---------------------------------------------------------------------------
# pure schedule generators
# ---------------------------------------------------------------------------

def schedule_no_pipelining(M: int, forward_only: bool = False) -> List[Task]:
    out = []
    for m in range(M):
        out.append(Task("F", m))
        if not forward_only:
            out.append(Task("B", m))
    return out


def schedule_1f1b(pp: int, rank: int, M: int, forward_only: bool = False) -> List[Task]:
    if forward_only:
        return [Task("F", m) for m in range(M)]
    warm = min(pp - rank - 1, M)
    out = [Task("F", m) for m in range(warm)]
    f, b = warm, 0
    for _ in range(M - warm):
        out.append(Task("F", f)); f += 1
        out.append(Task("B", b)); b += 1
    while b < M:
        out.append(Task("B", b)); b += 1
    return out


def get_schedule_table(M: int, C: int, G: int) -> List[Tuple[int, int]]:
    """Virtual-microbatch visiting order: for each group of G microbatches, for each chunk, for
    each microbatch of the group -> (microbatch, chunk)."""
    table = []
    for g0 in range(0, M, G):
        g1 = min(M, g0 + G)
        for c in range(C):
            for m in range(g0, g1):
                table.append((m, c))
    return table


def interleaved_warmup(pp: int, rank: int, M: int, C: int, G: int) -> int:
    total = M * C
    return min(total, (pp - rank - 1) * 2 + (C - 1) * G)


def validate_interleaved(pp: int, M: int, C: int, G: int) -> None:
    if not (pp <= G <= M):
        raise ValueError(f"need pp <= microbatch_group_size_per_vp_stage <= num_microbatches (pp={pp}, G={G}, M={M})")
    r = M % G
    if not (r == 0 or pp <= r < G):
        raise ValueError(f"num_microbatches % group_size must be 0 or in [pp, G): M={M}, G={G}, pp={pp}")


def schedule_interleaved(pp: int, rank: int, M: int, C: int, G: Optional[int] = None,
                         forward_only: bool = False) -> List[Task]:
    G = pp if G is None else G
    validate_interleaved(pp, M, C, G)
    table = get_schedule_table(M, C, G)
    total = len(table)
    if forward_only:
        return [Task("F", m, c) for m, c in table]
    warm = interleaved_warmup(pp, rank, M, C, G)
    out = [Task("F", *table[k]) for k in range(warm)]
    f, b = warm, 0
    for _ in range(total - warm):
        out.append(Task("F", *table[f])); f += 1
        m, c = table[b]
        out.append(Task("B", m, C - 1 - c)); b += 1
    while b < total:
        m, c = table[b]
        out.append(Task("B", m, C - 1 - c)); b += 1
    return out


# ---------------------------------------------------------------------------
# forward / backward steps
# ---------------------------------------------------------------------------

def deallocate_output_tensor(out: Optional[torch.Tensor], deallocate_pipeline_outputs: bool = False) -> None:
    """Pseudo-free an activation that was already sent downstream: only its ``grad_fn`` is
    needed for backward, so replace ``.data`` with a scalar."""
    if out is None or not deallocate_pipeline_outputs:
        return
    out.data = torch.empty((1,), device=out.device, dtype=out.dtype)


def custom_backward(output: torch.Tensor, grad_output: Optional[torch.Tensor]) -> None:
    """Backward that tolerates a deallocated ``output`` (shape check bypassed by calling the
    engine directly)."""
    if grad_output is None:
        grad_output = torch.ones_like(output, memory_format=torch.preserve_format)
    torch.autograd.Variable._execution_engine.run_backward(
        tensors=(output,), grad_tensors=(grad_output,), keep_graph=False, create_graph=False,
        inputs=tuple(), allow_unreachable=True, accumulate_grad=True)


def forward_step(forward_step_func, data_iterator, model, num_microbatches, input_tensor, forward_data_store,
                 config, collect_non_loss_data=False, checkpoint_activations_microbatch=None,
                 is_first_microbatch=False, current_microbatch=None, is_last_stage=None):
    if is_first_microbatch and hasattr(model, "set_is_first_microbatch"):
        model.set_is_first_microbatch()
    unwrap = model
    while hasattr(unwrap, "module") and not hasattr(unwrap, "set_input_tensor"):
        unwrap = unwrap.module
    model.set_input_tensor(input_tensor) if hasattr(model, "set_input_tensor") else unwrap.set_input_tensor(input_tensor)
    ctx = torch.autocast("cuda", dtype=config.autocast_dtype) if config.enable_autocast else contextlib.nullcontext()
    with _scope("forward", microbatch=current_microbatch), ctx:
        output_tensor, loss_func = forward_step_func(data_iterator, model)
    last = ps.is_pipeline_last_stage() if is_last_stage is None else is_last_stage
    num_tokens = torch.tensor(0, dtype=torch.int)
    if last:
        if not collect_non_loss_data:
            with _scope("loss"):
                outputs = loss_func(output_tensor)
            if len(outputs) == 3:
                output_tensor, num_tokens, loss_reduced = outputs
                if not config.calculate_per_token_loss:
                    output_tensor = output_tensor / num_tokens
                    output_tensor = output_tensor / num_microbatches
            else:
                output_tensor, loss_reduced = outputs
                output_tensor = output_tensor / num_microbatches
            forward_data_store.append(loss_reduced)
        else:
            forward_data_store.append(loss_func(output_tensor, non_loss_data=True))
    # Auxiliary losses (MoE load-balancing / z-loss, MTP) enter autograd through auto-scaler functions
    # whose backward scale must equal the main loss's: the fp16 loss scale, divided by the number
    # of microbatches unless the loss is per-token (reference schedules.py:317-339).
    n_experts = getattr(config, "num_moe_experts", None)
    mtp_layers = getattr(config, "mtp_num_layers", None)
    if n_experts is not None or mtp_layers:
        dev = output_tensor.device if torch.is_tensor(output_tensor) else None
        one = torch.ones(1, device=dev)
        scale = config.grad_scale_func(one) if config.grad_scale_func is not None else one
        if not config.calculate_per_token_loss:
            scale = scale / num_microbatches
        if n_experts is not None:
            from ..moe.moe_utils import MoEAuxLossAutoScaler

            MoEAuxLossAutoScaler.set_loss_scale(scale)
        if mtp_layers:
            from ..transformer.multi_token_prediction import MTPLossAutoScaler

            MTPLossAutoScaler.set_loss_scale(scale)
    return output_tensor, num_tokens


def backward_step(input_tensor, output_tensor, output_tensor_grad, model_type, config, microbatch=None):
    if input_tensor is not None and input_tensor.requires_grad:
        input_tensor.retain_grad()
    with _scope("backward", microbatch=microbatch):
        if output_tensor_grad is None and config.grad_scale_func is not None:
            output_tensor = config.grad_scale_func(output_tensor)
        if config.deallocate_pipeline_outputs:
            custom_backward(output_tensor, output_tensor_grad)
        else:
            torch.autograd.backward(output_tensor, grad_tensors=output_tensor_grad)
    return None if input_tensor is None else input_tensor.grad


def _backward_step_encdec(inputs, outputs, output_grads, config, microbatch=None):
    """Backward of one stage of a pipeline-split encoder-decoder model.  ``inputs`` / ``outputs`` / ``output_grads``
    are lists (or None on the first / last stage).  A decoder stage forwards the encoder output unchanged, so the
    gradient that comes back for it is added to this stage's own cross-attention gradient (reference
    ``schedules.backward_step``, encoder_and_decoder branch)."""
    inputs = [] if inputs is None else _as_list(inputs)
    outs = _as_list(outputs)
    grads = [None] * len(outs) if output_grads is None else _as_list(output_grads)
    for x in inputs:
        if x is not None and x.requires_grad:
            x.retain_grad()
    with _scope("backward", microbatch=microbatch):
        head = outs[0]
        if grads[0] is None and config.grad_scale_func is not None:
            head = config.grad_scale_func(head)
        torch.autograd.backward(head, grad_tensors=grads[0])
    in_grads = [None if x is None else x.grad for x in inputs]
    if len(grads) > 1 and grads[1] is not None and in_grads:
        in_grads[-1] = grads[1] if in_grads[-1] is None else in_grads[-1].add_(grads[1])
    return in_grads


def _unwrap_chunk(m):
    while hasattr(m, "module"):
        m = m.module
    return m


# ---------------------------------------------------------------------------
# the executor
# ---------------------------------------------------------------------------

def _get_config(model):
    m = model
    while hasattr(m, "module"):
        m = m.module
    return m.config


def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


def get_tensor_shapes(seq_length, micro_batch_size, config) -> Tuple[int, int, int]:
    s = seq_length // max(1, config.context_parallel_size)
    if config.sequence_parallel:
        s = s // config.tensor_model_parallel_size
    return (s, micro_batch_size, config.hidden_size)


def _encdec_shapes(rank: int, seq_length: int, decoder_seq_length: int, micro_batch_size: int, config) -> List[Tuple]:
    """Activations crossing the boundary AFTER pipeline stage ``rank`` of an encoder-decoder model (reference
    ``schedules.get_tensor_shapes``): encoder stages pass the encoder stream; decoder stages pass the decoder stream
    AND the (finished) encoder output, which every decoder stage needs for cross attention."""
    enc = get_tensor_shapes(seq_length, micro_batch_size, config)
    if ps.is_inside_encoder(rank):
        return [enc]
    return [get_tensor_shapes(decoder_seq_length, micro_batch_size, config), enc]


def run_tasks(tasks: List[Task], *, forward_step_func, data_iterator, model: List, num_microbatches: int,
              seq_length: int, micro_batch_size: int, forward_only: bool, collect_non_loss_data: bool,
              first_val_step=None, on_task: Optional[Callable] = None, decoder_seq_length: Optional[int] = None):
    """Execute a task list on this pipeline rank."""
    chunks = _as_list(model)
    C = len(chunks)
    iters = _as_list(data_iterator) if data_iterator is not None else [None] * C
    if len(iters) != C:
        iters = iters * C
    config = _get_config(chunks[0])
    pp = ps.get_pipeline_model_parallel_world_size()
    rank = ps.get_pipeline_model_parallel_rank()
    shape = get_tensor_shapes(seq_length, micro_batch_size, config)
    encdec = pp > 1 and ps.get_pipeline_model_parallel_split_rank() is not None \
        and str(getattr(chunks[0], "model_type", None) or getattr(_unwrap_chunk(chunks[0]), "model_type", None)) == "encoder_and_decoder"
    if encdec:
        if C > 1:
            raise ValueError("interleaved schedule is not supported for encoder-decoder models")
        recv_shapes = _encdec_shapes(rank - 1, seq_length, decoder_seq_length or seq_length, micro_batch_size, config) if rank > 0 else []
        send_shapes = _encdec_shapes(rank, seq_length, decoder_seq_length or seq_length, micro_batch_size, config)
    dtype = config.pipeline_dtype or config.params_dtype
    prev_r = ps.get_pipeline_model_parallel_prev_rank() if pp > 1 else None
    next_r = ps.get_pipeline_model_parallel_next_rank() if pp > 1 else None
    forward_data_store: List = []
    total_num_tokens = torch.zeros((), dtype=torch.int, device=next(chunks[0].parameters()).device)
    stash: Dict[Tuple[int, int], Tuple] = {}
    last_b = {c: max((t.microbatch for t in tasks if t.kind == "B" and t.chunk == c), default=-1) for c in range(C)}
    first_f_done = [False] * C
    vpp = C if C > 1 else None

    def is_first(c):
        return rank == 0 and c == 0

    def is_last(c):
        return rank == pp - 1 and c == C - 1

    for chunk in chunks:
        if hasattr(chunk, "is_last_microbatch"):
            chunk.is_last_microbatch = False
    for t in tasks:
        c = t.chunk
        if vpp:
            ps.set_virtual_pipeline_model_parallel_rank(c)
        if on_task is not None:
            on_task(t)
        if t.kind == "F":
            inp = None
            if encdec and not is_first(c):
                inp = [p2p.recv(sh, dtype, prev_r, "fwd", "recv-forward", requires_grad=not forward_only) for sh in recv_shapes]
            elif pp > 1 and not is_first(c):
                inp = p2p.recv(shape, dtype, prev_r, "fwd", "recv-forward", requires_grad=not forward_only)
            ctx = torch.no_grad() if forward_only else contextlib.nullcontext()
            with ctx:
                out, ntok = forward_step(forward_step_func, iters[c], chunks[c], num_microbatches, inp,
                                         forward_data_store, config, collect_non_loss_data,
                                         is_first_microbatch=not first_f_done[c], current_microbatch=t.microbatch,
                                         is_last_stage=is_last(c))
            first_f_done[c] = True
            total_num_tokens += ntok.to(total_num_tokens.device) if torch.is_tensor(ntok) else int(ntok)
            if encdec and not is_last(c):
                if ps.is_inside_decoder(rank):
                    out = [out, inp[-1]]                 # hand the encoder output on to the next decoder stage
                for o in _as_list(out):
                    p2p.send(o, next_r, "fwd", "send-forward")
            elif pp > 1 and not is_last(c):
                p2p.send(out, next_r, "fwd", "send-forward")
                if not forward_only:
                    deallocate_output_tensor(out, config.deallocate_pipeline_outputs)
            if not forward_only:
                stash[(t.microbatch, c)] = (inp, out)
        else:
            inp, out = stash.pop((t.microbatch, c))
            grad = None
            if encdec and not is_last(c):
                grad = [p2p.recv(sh, dtype, next_r, "bwd", "recv-backward") for sh in send_shapes]
            elif pp > 1 and not is_last(c):
                grad = p2p.recv(shape, dtype, next_r, "bwd", "recv-backward")
            if hasattr(chunks[c], "is_last_microbatch"):
                chunks[c].is_last_microbatch = (t.microbatch == last_b[c])
            if encdec:
                in_grad = _backward_step_encdec(inp, out, grad, config, microbatch=t.microbatch)
                for g_ in (in_grad if not is_first(c) else []):
                    p2p.send(g_, prev_r, "bwd", "send-backward")
                continue
            in_grad = backward_step(inp, out, grad, None, config, microbatch=t.microbatch)
            if pp > 1 and not is_first(c):
                p2p.send(in_grad, prev_r, "bwd", "send-backward")
    if pp > 1:
        p2p.flush()
    if vpp:
        ps.set_virtual_pipeline_model_parallel_rank(0)
    for chunk in chunks:
        if hasattr(chunk, "is_last_microbatch"):
            chunk.is_last_microbatch = True
    if not forward_only and config.finalize_model_grads_func is not None:
        config.finalize_model_grads_func(chunks, total_num_tokens if config.calculate_per_token_loss else None)
    return forward_data_store


def forward_backward_no_pipelining(*, forward_step_func, data_iterator, model, num_microbatches, seq_length,
                                   micro_batch_size, decoder_seq_length=None, forward_only=False,
                                   collect_non_loss_data=False, first_val_step=None):
    return run_tasks(schedule_no_pipelining(num_microbatches, forward_only), forward_step_func=forward_step_func,
                     data_iterator=data_iterator, model=model, num_microbatches=num_microbatches,
                     seq_length=seq_length, micro_batch_size=micro_batch_size, forward_only=forward_only,
                     collect_non_loss_data=collect_non_loss_data, first_val_step=first_val_step,
                     decoder_seq_length=decoder_seq_length)


def forward_backward_pipelining_without_interleaving(*, forward_step_func, data_iterator, model, num_microbatches,
                                                     seq_length, micro_batch_size, decoder_seq_length=None,
                                                     forward_only=False, collect_non_loss_data=False,
                                                     first_val_step=None):
    tasks = schedule_1f1b(ps.get_pipeline_model_parallel_world_size(), ps.get_pipeline_model_parallel_rank(),
                          num_microbatches, forward_only)
    return run_tasks(tasks, forward_step_func=forward_step_func, data_iterator=data_iterator, model=model,
                     num_microbatches=num_microbatches, seq_length=seq_length, micro_batch_size=micro_batch_size,
                     forward_only=forward_only, collect_non_loss_data=collect_non_loss_data,
                     first_val_step=first_val_step, decoder_seq_length=decoder_seq_length)


def forward_backward_pipelining_with_interleaving(*, forward_step_func, data_iterator, model, num_microbatches,
                                                  seq_length, micro_batch_size, decoder_seq_length=None,
                                                  forward_only=False, collect_non_loss_data=False,
                                                  first_val_step=None):
    chunks = _as_list(model)
    config = _get_config(chunks[0])
    tasks = schedule_interleaved(ps.get_pipeline_model_parallel_world_size(),
                                 ps.get_pipeline_model_parallel_rank(), num_microbatches, len(chunks),
                                 config.microbatch_group_size_per_vp_stage, forward_only)
    return run_tasks(tasks, forward_step_func=forward_step_func, data_iterator=data_iterator, model=model,
                     num_microbatches=num_microbatches, seq_length=seq_length, micro_batch_size=micro_batch_size,
                     forward_only=forward_only, collect_non_loss_data=collect_non_loss_data,
                     first_val_step=first_val_step)


_SCHEDULE_OVERRIDE: Optional[Callable] = None


def set_schedule_override(fn: Optional[Callable]) -> None:
    """MegaFBD / MegaDPP install their schedule here (``--forward-backward-disaggregating``,
    ``--use-dpp``) without the core importing the training harness."""
    global _SCHEDULE_OVERRIDE
    _SCHEDULE_OVERRIDE = fn


def get_forward_backward_func():
    if _SCHEDULE_OVERRIDE is not None:
        return _SCHEDULE_OVERRIDE
    pp = ps.get_pipeline_model_parallel_world_size()
    if pp > 1:
        if ps.get_virtual_pipeline_model_parallel_world_size() is not None:
            return forward_backward_pipelining_with_interleaving
        return forward_backward_pipelining_without_interleaving
    return forward_backward_no_pipelining


# -- helper names of the reference's ``pipeline_parallel/schedules.py`` ---------------------------------------------------------------
def get_pp_rank_microbatches(num_microbatches: int, num_model_chunks: int, microbatch_group_size_per_vp_stage: Optional[int],
                             forward_only: bool = False):
    """-> ``(total virtual microbatches, all in warm-up?, warm-up count, remaining count)`` for this pipeline rank."""
    pp, rank = ps.get_pipeline_model_parallel_world_size(), ps.get_pipeline_model_parallel_rank()
    total = num_microbatches * num_model_chunks
    if forward_only:
        warm = total
    elif num_model_chunks > 1:
        warm = interleaved_warmup(pp, rank, num_microbatches, num_model_chunks, microbatch_group_size_per_vp_stage or pp)
    else:
        warm = min(pp - rank - 1, total)
    return total, warm >= total, warm, total - warm


def convert_schedule_table_to_order(num_warmup_microbatches: int, num_model_chunks: int, schedule_table) -> List[int]:
    """Execution order as signed chunk numbers: ``+(chunk + 1)`` = a forward of that chunk, ``-(chunk + 1)`` = a backward
    (backwards visit the chunks in reverse: chunk ``C - 1 - c``); warm-up forwards, then 1F1B pairs, then the cool-down backwards."""
    _, chunk_ids = zip(*schedule_table)
    fwd = [c + 1 for c in chunk_ids]
    bwd = [c - num_model_chunks for c in chunk_ids]           # -(C - c): chunk C-1-c, negated and 1-based
    order = fwd[:num_warmup_microbatches]
    for i in range(num_warmup_microbatches, len(fwd)):
        order.append(fwd[i])
        order.append(bwd[i - num_warmup_microbatches])
    if num_warmup_microbatches > 0:
        order.extend(bwd[-num_warmup_microbatches:])
    return order


def check_first_val_step(first_val_step: Optional[bool], forward_only: bool, cond: bool) -> bool:
    """fp8 weight caching: recompute cached weights only on the first validation microbatch."""
    if first_val_step is not None and forward_only:
        return first_val_step and cond
    return cond


def set_current_microbatch(model, microbatch_id: int) -> None:
    """Tell the layers which microbatch runs (CUDA-graph / fp8 weight-cache bookkeeping keyed by microbatch)."""
    core = _unwrap_chunk(model)
    decoder = getattr(core, "decoder", None)
    for layer in getattr(decoder, "layers", []) or []:
        layer.current_microbatch = microbatch_id


def clear_embedding_activation_buffer(config, model):
    """Drop the output layer's stashed activations of ``defer_embedding_wgrad_compute`` (last stage only)."""
    if ps.is_pipeline_last_stage(ignore_virtual=True) and getattr(config, "defer_embedding_wgrad_compute", False):
        m = _unwrap_chunk(model[-1] if isinstance(model, list) else model)
        if getattr(m, "share_embeddings_and_output_weights", False) or hasattr(m, "output_layer"):
            buf = getattr(m, "embedding_activation_buffer", None)
            if buf is not None:
                buf.clear()
            return m
    return None


def finish_embedding_wgrad_compute(config, embedding_module):
    """Run the deferred output-layer weight-gradient GEMMs (see ``core_api.utils.drain_embedding_wgrad_compute``)."""
    if ps.is_pipeline_last_stage(ignore_virtual=True) and getattr(config, "defer_embedding_wgrad_compute", False) \
            and embedding_module is not None:
        from ..core_api.utils import drain_embedding_wgrad_compute

        weight = embedding_module.shared_embedding_or_output_weight() \
            if getattr(embedding_module, "share_embeddings_and_output_weights", False) else embedding_module.output_layer.weight
        drain_embedding_wgrad_compute(config, getattr(embedding_module, "embedding_activation_buffer", []),
                                      getattr(embedding_module, "grad_output_buffer", []), weight)


def recv_forward(tensor_shapes, config):
    shapes = tensor_shapes if isinstance(tensor_shapes, list) else [tensor_shapes]
    return [None if s is None else p2p.recv_forward(s, config) for s in shapes]


def recv_backward(tensor_shapes, config):
    shapes = tensor_shapes if isinstance(tensor_shapes, list) else [tensor_shapes]
    return [None if s is None else p2p.recv_backward(s, config) for s in shapes]


def send_forward(output_tensors, tensor_shapes, config):
    outs = output_tensors if isinstance(output_tensors, list) else [output_tensors]
    for o in outs:
        if o is not None:
            p2p.send_forward(o, config)


def send_backward(input_tensor_grads, tensor_shapes, config):
    grads = input_tensor_grads if isinstance(input_tensor_grads, list) else [input_tensor_grads]
    for g in grads:
        if g is not None:
            p2p.send_backward(g, config)


def send_corresponding_forward(tensor, config=None):
    """MegaFBD: a forward rank hands a stage input to its backward twin."""
    return p2p.send_corresponding_forward(tensor, config)


def recv_corresponding_forward(tensor_shape, dtype):
    return p2p.recv_corresponding_forward(tensor_shape, dtype)


def send_forward_recv_backward(output_tensors, tensor_shapes, config):
    outs = output_tensors if isinstance(output_tensors, list) else [output_tensors]
    shapes = tensor_shapes if isinstance(tensor_shapes, list) else [tensor_shapes]
    return [None if s is None else p2p.send_forward_recv_backward(o, s, config) for o, s in zip(outs, shapes)]


def send_backward_recv_forward(input_tensor_grads, tensor_shapes, config):
    grads = input_tensor_grads if isinstance(input_tensor_grads, list) else [input_tensor_grads]
    shapes = tensor_shapes if isinstance(tensor_shapes, list) else [tensor_shapes]
    return [None if s is None else p2p.send_backward_recv_forward(g, s, config) for g, s in zip(grads, shapes)]


def forward_step_no_grad(*args, **kwargs):
    """``forward_step`` under ``torch.no_grad`` (MegaFBD forward ranks, evaluation)."""
    with torch.no_grad():
        return forward_step(*args, **kwargs)
