"""Shared (always-on) expert of a MoE layer (reference ``transformer/moe/shared_experts.py`` ``SharedExpertMLP``): a dense
gated MLP of width ``moe_shared_expert_intermediate_size`` applied to every token and added to the routed experts' output;
optionally scaled by a learned sigmoid gate (Qwen2-MoE style, ``gate=True`` in the layer spec).

``--moe-shared-expert-overlap``: the reference slices the shared expert's GEMMs between the dispatcher's all-to-alls on one
stream and orders them with ``CUDA_DEVICE_MAX_CONNECTIONS=1``.  On B200 the routed path is a chain of small kernels (route,
dispatch over NVLink, grouped GEMM, combine) that leaves SMs idle, so the shared expert simply runs on a second stream:
``start(x)`` forks it before the dispatch, ``finish()`` joins it before the add.  Autograd replays the same fork / join in
backward because every op remembers the stream it ran on."""
from __future__ import annotations

import copy

import torch
import torch.nn as nn

from ..transformer.mlp import MLP


class SharedExpertMLP(MLP):
    def __init__(self, config, submodules, gate: bool = False):
        cfg = copy.copy(config)
        cfg.ffn_hidden_size = config.moe_shared_expert_intermediate_size
        super().__init__(cfg, submodules)
        self.use_shared_expert_gate = bool(gate)
        if gate:
            self.gate_weight = nn.Parameter(torch.empty(1, config.hidden_size, dtype=config.params_dtype,
                                                        device=self.linear_fc2.weight.device))
            if config.perform_initialization:
                config.init_method(self.gate_weight)
            self.gate_weight.sequence_parallel = config.sequence_parallel      # replicated over TP: grads all-reduced
        else:
            self.gate_weight = None
        self.overlap = bool(config.moe_shared_expert_overlap)
        self._stream = None
        self._pending = None

    def forward(self, hidden_states):
        out, bias = super().forward(hidden_states)
        if bias is not None:
            out = out + bias
        if self.gate_weight is not None:
            logits = torch.nn.functional.linear(hidden_states, self.gate_weight.to(hidden_states.dtype))
            out = out * torch.sigmoid(logits.float()).to(out.dtype)
        return out

    # ---- overlap with the routed path -------------------------------------------------------------------------------------
    def start(self, hidden_states) -> None:
        """Launch the shared expert; with overlap on a CUDA device it runs on a side stream while the caller dispatches."""
        if not (self.overlap and hidden_states.is_cuda):
            self._pending = (None, hidden_states)
            return
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=hidden_states.device)
        cur = torch.cuda.current_stream(hidden_states.device)
        self._stream.wait_stream(cur)
        with torch.cuda.stream(self._stream):
            out = self.forward(hidden_states)
        hidden_states.record_stream(self._stream)
        self._pending = (out, None)

    def finish(self) -> torch.Tensor:
        out, x = self._pending
        self._pending = None
        if out is None:
            return self.forward(x)
        cur = torch.cuda.current_stream(out.device)
        cur.wait_stream(self._stream)
        out.record_stream(cur)
        return out
