"""CUDA-graph capture / replay of transformer layers and of the decode step.

Parity: reference ``transformer/cuda_graphs.py`` (``CudaGraphManager :745`` -- per-layer capture of
forward and backward after ``cuda_graph_warmup_steps`` eager iterations, static input/output
buffers, replay afterwards; the TE ``make_graphed_callables`` path of ``training.py:454-630``;
flags ``--enable-cuda-graph`` / ``--cuda-graph-warmup-steps`` / ``cuda_graph_scope``).

Launch-bound regimes (small micro-batches under TP/PP, single-token decode) are where this pays on
B200: a transformer layer is ~35 kernels of 5-40 us each.  Every kernel of this framework launches
on the current stream with no host synchronisation, so a layer is capturable as is; the two
exceptions are refused explicitly: the fused TP AG/RS GEMMs (their flag epochs are host-side
kernel arguments) and MoE layers with data-dependent shapes.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn as nn


def is_graph_capturable(layer: nn.Module) -> Tuple[bool, str]:
    from ..parallel import tp_layers

    if not torch.cuda.is_available():
        return False, "no CUDA device"
    if getattr(tp_layers, "_FUSED_COMM", None) is not None:
        return False, "fused TP comm kernels carry host-side epochs"
    for m in layer.modules():
        if m.__class__.__name__ in ("MoELayer",):
            return False, "MoE layer has data-dependent shapes"
    return True, ""


class _GraphedLayer(nn.Module):
    """A layer whose forward+backward run as two CUDA graphs (via ``make_graphed_callables``)."""

    def __init__(self, layer: nn.Module, sample_args: Tuple[torch.Tensor, ...], pool=None, num_warmup_iters: int = 3):
        super().__init__()
        self.layer = layer
        self._graphed = torch.cuda.make_graphed_callables(_TensorOnly(layer), sample_args,
                                                          num_warmup_iters=num_warmup_iters, pool=pool)

    def forward(self, *args):
        return self._graphed(*args)


class _TensorOnly(nn.Module):
    """Adapts ``layer(hidden, ...) -> (hidden, context)`` to tensors-in / tensor-out."""

    def __init__(self, layer):
        super().__init__()
        self.layer = layer

    def forward(self, hidden_states, rotary_pos_emb=None):
        out = self.layer(hidden_states, rotary_pos_emb=rotary_pos_emb)
        return out[0] if isinstance(out, tuple) else out


class CudaGraphManager:
    """Eager for ``warmup_steps`` calls, then capture once per (shape, dtype, training) and replay."""

    def __init__(self, layer: nn.Module, warmup_steps: int = 3, share_pool: bool = False):
        self.layer = layer
        self.warmup_steps = warmup_steps
        self.calls = 0
        self.graphs: Dict[tuple, Callable] = {}
        self.refused: Optional[str] = None
        self.pool = torch.cuda.graph_pool_handle() if (share_pool and torch.cuda.is_available()) else None
        self.replays = 0

    def _key(self, args):
        return tuple((tuple(a.shape), a.dtype, a.requires_grad) for a in args) + (self.layer.training,)

    def __call__(self, hidden_states: torch.Tensor, *extra: torch.Tensor):
        args = (hidden_states,) + tuple(extra)
        self.calls += 1
        if self.refused or self.calls <= self.warmup_steps or not hidden_states.is_cuda:
            return _TensorOnly(self.layer)(*args)
        key = self._key(args)
        g = self.graphs.get(key)
        if g is None:
            ok, why = is_graph_capturable(self.layer)
            if not ok:
                self.refused = why
                return _TensorOnly(self.layer)(*args)
            sample = tuple(torch.zeros_like(a).requires_grad_(a.requires_grad) for a in args)
            # capture at a quiescent point: drop dead autograd graphs / cross-stream blocks first so that no
            # deferred free lands inside the capture (it would touch the legacy stream and invalidate it)
            import gc

            gc.collect()
            torch.cuda.synchronize()
            g = torch.cuda.make_graphed_callables(_TensorOnly(self.layer), sample, num_warmup_iters=3, pool=self.pool)
            self.graphs[key] = g
        self.replays += 1
        return g(*args)


class DecodeStepGraph:
    """Whole-model single-token decode step captured as one graph (static token / position buffers,
    KV cache updated in place by the captured kernels).  ``step(tokens, positions)`` replays."""

    def __init__(self, model: nn.Module, inference_context, batch_size: int, device=None):
        self.model, self.ctx = model, inference_context
        dev = device or torch.device("cuda", torch.cuda.current_device())
        self.tokens = torch.zeros(batch_size, 1, dtype=torch.int64, device=dev)
        self.positions = torch.zeros(batch_size, 1, dtype=torch.int64, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.logits: Optional[torch.Tensor] = None
        self.captured_offset: Optional[int] = None

    def capture(self):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(2):
                self.model(self.tokens, self.positions, None, inference_context=self.ctx)
        torch.cuda.current_stream().wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.logits = self.model(self.tokens, self.positions, None, inference_context=self.ctx)
        self.captured_offset = self.ctx.sequence_len_offset

    def step(self, tokens: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        """Replay is only valid at the sequence offset the graph was captured for (the KV-cache
        slice bounds are baked in); callers re-capture per offset bucket or fall back to eager."""
        if self.graph is None or self.captured_offset != self.ctx.sequence_len_offset:
            self.capture()
        self.tokens.copy_(tokens)
        self.positions.copy_(positions)
        self.graph.replay()
        return self.logits


def graph_transformer_block(block: nn.Module, warmup_steps: int = 3) -> List[CudaGraphManager]:
    """Attach a ``CudaGraphManager`` to every layer of a ``TransformerBlock`` (``--enable-cuda-graph``)."""
    managers = []
    for layer in block.layers:
        m = CudaGraphManager(layer, warmup_steps)
        layer.cuda_graph_manager = m
        managers.append(m)
    return managers


def capture_block_upfront(block: nn.Module, hidden_states: torch.Tensor, *extra: torch.Tensor) -> None:
    """Capture every layer of the block with zero-filled sample inputs *before* any of them runs in this
    forward pass (all layers of a block see the same shapes), instead of lazily in the middle of it."""
    for layer in block.layers:
        m = layer.cuda_graph_manager
        args = (hidden_states,) + tuple(extra)
        if m.refused is None and m._key(args) not in m.graphs and m.calls >= m.warmup_steps:
            m.calls += 1
            saved = m.calls
            m(*(torch.zeros_like(a).requires_grad_(a.requires_grad) for a in args))   # triggers the capture
            m.calls, m.replays = saved, max(0, m.replays - 1)


def is_graph_capturing() -> bool:
    """True while a CUDA graph is being captured on the current stream."""
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def create_cudagraphs() -> None:
    """The reference records graph-able calls first and captures them all here.  ``CudaGraphManager`` captures a layer the first
    time it is called after its warm-up steps (or up front through ``capture_block_upfront``), so nothing is pending."""
    return None
