"""Pipeline point-to-point transport: unidirectional NCCL/gloo channels.

Capability parity with reference ``core/pipeline_parallel/p2p_communication.py`` (send/recv of
``[s/(tp if SP), b, h]`` activations and gradients between adjacent stages, optional shape
handshake, MegaScan ``send-*``/``recv-*`` scopes with ``data`` bytes, FBD twin transfers
:723-756) -- but instead of ``batch_isend_irecv`` with even/odd ordering tricks, every directed
edge of the pipeline ring gets its own 2-rank communicator ("channel").  A channel carries
messages in one direction only, so a blocked send can never sit in front of a receive on the
same NCCL stream: any schedule whose data-flow graph is acyclic -- 1F1B, interleaved, the
MegaFBD forward/backward split, MegaDPP's dynamic order -- is deadlock-free by construction,
without the runtime coordinator the reference needs for FBD (``Controller.py``).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import state as ps
from ..scan.tracer import scope as _scope

_CHANNELS: Dict[Tuple[int, int, str], object] = {}
_PENDING: Dict[Tuple[int, int, str], list] = {}          # channel -> [(work handle, tensor kept alive)]
_MAX_IN_FLIGHT = 8


def build_channels() -> None:
    """Create one group per directed neighbour edge and traffic class; collective over WORLD."""
    _CHANNELS.clear()
    if not (dist.is_available() and dist.is_initialized()) or not ps.model_parallel_is_initialized():
        return
    topo = ps.get_topology()
    edges: List[Tuple[int, int, str]] = []
    for g in topo.groups("pp"):
        n = len(g)
        if n < 2:
            continue
        for i in range(n):
            a, b = g[i], g[(i + 1) % n]
            edges.append((a, b, "fwd"))   # activations a -> b
            edges.append((b, a, "bwd"))   # gradients  b -> a
    if topo.fbd:
        for g in topo.groups("fb"):
            for b in g[1:]:
                edges.append((g[0], b, "fbd"))    # forward rank -> backward twin(s): stage inputs
                if topo.fbd_collapse:
                    edges.append((b, g[0], "fbw"))  # backward TP rank -> un-sharded forward rank: updated weight shards
    seen = set()
    for a, b, kind in edges:
        if (a, b, kind) in seen:
            continue
        seen.add((a, b, kind))
        grp = dist.new_group(ranks=sorted({a, b}))
        _CHANNELS[(a, b, kind)] = grp


def destroy_channels() -> None:
    _CHANNELS.clear()
    _PENDING.clear()


def _channel(src: int, dst: int, kind: str):
    key = (src, dst, kind)
    if key not in _CHANNELS:
        raise RuntimeError(f"no p2p channel {key}; call build_channels() after initialize_model_parallel()")
    return _CHANNELS[key]


def send(tensor: torch.Tensor, dst: int, kind: str, name: str = "send") -> None:
    """Asynchronous send.  Messages of one channel stay ordered; a short queue of in-flight sends is kept per channel
    (completed ones are retired, the oldest is waited for only when ``_MAX_IN_FLIGHT`` is exceeded) -- a stage that
    sends two tensors per boundary (encoder-decoder split) must not block on its first send while its peer is
    itself sending."""
    me = dist.get_rank()
    key = (me, dst, kind)
    q = _PENDING.setdefault(key, [])
    while q and q[0][0].is_completed():
        q.pop(0)
    while len(q) >= _MAX_IN_FLIGHT:
        q.pop(0)[0].wait()
    # a detached alias: the caller may pseudo-free ``tensor`` (swap its .data) right after this
    # returns, while the asynchronous send still needs the storage
    t = tensor.detach().contiguous()
    if t is tensor:
        t = tensor.detach()
    with _scope(name, data=t.numel() * t.element_size(), peer=dst):
        h = dist.isend(t, dst=dst, group=_channel(me, dst, kind))
    q.append((h, t))


def recv(shape, dtype, src: int, kind: str, name: str = "recv", device=None, requires_grad: bool = False):
    me = dist.get_rank()
    device = device or (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu")
    buf = torch.empty(shape, dtype=dtype, device=device, requires_grad=False)
    with _scope(name, data=buf.numel() * buf.element_size(), peer=src):
        dist.irecv(buf, src=src, group=_channel(src, me, kind)).wait()
    if requires_grad:
        buf.requires_grad_(True)
    return buf


def flush() -> None:
    for key in list(_PENDING):
        for h, _ in _PENDING.pop(key):
            h.wait()


# ---- Megatron-named helpers (non-interleaved convenience API) ---------------------------

def _shape(config, tensor_shape):
    return tuple(tensor_shape)


def recv_forward(tensor_shape, config):
    if ps.is_pipeline_first_stage(ignore_virtual=True):
        return None
    return recv(tensor_shape, config.pipeline_dtype, ps.get_pipeline_model_parallel_prev_rank(), "fwd",
                "recv-forward", requires_grad=True)


def recv_backward(tensor_shape, config):
    if ps.is_pipeline_last_stage(ignore_virtual=True):
        return None
    return recv(tensor_shape, config.pipeline_dtype, ps.get_pipeline_model_parallel_next_rank(), "bwd",
                "recv-backward")


def send_forward(output_tensor, config):
    if not ps.is_pipeline_last_stage(ignore_virtual=True):
        send(output_tensor, ps.get_pipeline_model_parallel_next_rank(), "fwd", "send-forward")


def send_backward(input_tensor_grad, config):
    if not ps.is_pipeline_first_stage(ignore_virtual=True):
        send(input_tensor_grad, ps.get_pipeline_model_parallel_prev_rank(), "bwd", "send-backward")


def send_forward_recv_backward(output_tensor, tensor_shape, config):
    send_forward(output_tensor, config)
    return recv_backward(tensor_shape, config)


def send_backward_recv_forward(input_tensor_grad, tensor_shape, config):
    send_backward(input_tensor_grad, config)
    return recv_forward(tensor_shape, config)


def send_corresponding_forward(tensor, config=None):
    """MegaFBD: forward rank ships a stage input to its backward twin.  With the forward TP collapsed the one forward
    rank feeds all tp backward ranks of its block: each gets the full activation, or -- under sequence parallelism --
    its own ``s / tp`` slice of the sequence dimension."""
    if not ps.fbd_forward_tp_collapsed():
        send(tensor, ps.get_twin_rank(), "fbd", "fbd-send")
        return
    twins = ps.get_fbd_block_ranks()[1:]
    sp = bool(config is not None and getattr(config, "sequence_parallel", False))
    n = tensor.shape[0] // len(twins) if sp else 0
    for t, dst in enumerate(twins):
        send(tensor[t * n:(t + 1) * n] if sp else tensor, dst, "fbd", "fbd-send")


def recv_corresponding_forward(tensor_shape, dtype):
    src = ps.get_fbd_block_ranks()[0] if ps.fbd_forward_tp_collapsed() else ps.get_twin_rank()
    return recv(tensor_shape, dtype, src, "fbd", "fbd-recv", requires_grad=True)


def communicate_shapes(tensor: Optional[torch.Tensor], src_or_dst: int, kind: str, sending: bool):
    """3-int64 shape handshake for ``variable_seq_lengths`` (reference :27-122)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu"
    if sending:
        s = torch.tensor(list(tensor.shape) + [0] * (3 - tensor.dim()), dtype=torch.int64, device=dev)
        send(s, src_or_dst, kind, "send-shape")
        return tuple(tensor.shape)
    s = recv((3,), torch.int64, src_or_dst, kind, "recv-shape", device=dev)
    return tuple(int(v) for v in s.tolist() if v > 0)


# -- combined calls of the reference's p2p_communication.py (interleaved schedule steady state) ------------------------------
def send_forward_recv_forward(output_tensor, recv_prev: bool, tensor_shape, config, overlap_p2p_comm: bool = False):
    """Send this stage's output downstream and (optionally) receive the next input from upstream."""
    if output_tensor is not None and not ps.is_pipeline_last_stage(ignore_virtual=True):
        send_forward(output_tensor, config)
    got = recv_forward(tensor_shape, config) if (recv_prev and not ps.is_pipeline_first_stage(ignore_virtual=True)) else None
    return (got, []) if overlap_p2p_comm else got


def send_backward_recv_backward(input_tensor_grad, recv_next: bool, tensor_shape, config, overlap_p2p_comm: bool = False):
    if input_tensor_grad is not None and not ps.is_pipeline_first_stage(ignore_virtual=True):
        send_backward(input_tensor_grad, config)
    got = recv_backward(tensor_shape, config) if (recv_next and not ps.is_pipeline_last_stage(ignore_virtual=True)) else None
    return (got, []) if overlap_p2p_comm else got


def send_forward_backward_recv_forward_backward(output_tensor, input_tensor_grad, recv_prev: bool, recv_next: bool, tensor_shape, config):
    """Both directions at once -> ``(input_tensor, output_tensor_grad)``.  Sends are posted before the receives on the
    unidirectional channels, so two neighbouring stages calling this at the same time cannot deadlock."""
    if output_tensor is not None and not ps.is_pipeline_last_stage(ignore_virtual=True):
        send_forward(output_tensor, config)
    if input_tensor_grad is not None and not ps.is_pipeline_first_stage(ignore_virtual=True):
        send_backward(input_tensor_grad, config)
    inp = recv_forward(tensor_shape, config) if (recv_prev and not ps.is_pipeline_first_stage(ignore_virtual=True)) else None
    grad = recv_backward(tensor_shape, config) if (recv_next and not ps.is_pipeline_last_stage(ignore_virtual=True)) else None
    return inp, grad
