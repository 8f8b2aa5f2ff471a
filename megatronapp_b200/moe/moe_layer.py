"""Mixture-of-Experts layer: router -> dispatch -> experts (+ shared experts) -> combine.

Parity: reference ``transformer/moe/moe_layer.py`` (``MoELayer``, ``MoESubmodules``) and
``shared_experts.py``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Union

import torch

from ..parallel import state as ps
from ..transformer.mlp import MLP
from ..transformer.module import MegatronModule
from ..transformer.spec_utils import ModuleSpec, build_module
from .router import TopKRouter
from .token_dispatcher import (MoEAllGatherTokenDispatcher, MoEAlltoAllTokenDispatcher, MoEArenaDispatcher, MoEFusedDispatcher,
                               MoEPeerScatterDispatcher)


@dataclass
class MoESubmodules:
    experts: Union[ModuleSpec, type] = None
    shared_experts: Union[ModuleSpec, type] = None


class MoELayer(MegatronModule):
    def __init__(self, config, submodules: MoESubmodules = None, layer_number: int = None):
        super().__init__(config)
        self.layer_number = layer_number
        ep = ps.get_expert_model_parallel_world_size()
        if config.num_moe_experts % ep != 0:
            raise ValueError("num_moe_experts must be divisible by the expert parallel size")
        self.num_local_experts = config.num_moe_experts // ep
        off = ps.get_expert_model_parallel_rank() * self.num_local_experts
        self.local_expert_indices = [off + i for i in range(self.num_local_experts)]
        self.router = TopKRouter(config)
        kind = config.moe_token_dispatcher_type
        if kind == "allgather":
            self.token_dispatcher = MoEAllGatherTokenDispatcher(self.num_local_experts, self.local_expert_indices, config)
        elif kind in ("alltoall", "alltoall_seq"):
            self.token_dispatcher = MoEAlltoAllTokenDispatcher(self.num_local_experts, self.local_expert_indices, config)
        elif kind in ("flex", "fused") and ep == 1:
            self.token_dispatcher = MoEFusedDispatcher(self.num_local_experts, self.local_expert_indices, config)
        elif kind in ("flex", "fused") and torch.cuda.is_available() and not config.use_cpu_initialization:
            self.token_dispatcher = MoEArenaDispatcher(self.num_local_experts, self.local_expert_indices, config)
        elif kind == "flex":
            self.token_dispatcher = MoEPeerScatterDispatcher(self.num_local_experts, self.local_expert_indices, config)
        else:
            raise ValueError(f"unsupported token dispatcher type {kind}")
        self.experts = build_module(submodules.experts, self.num_local_experts, config)
        self.shared_experts = None
        if config.moe_shared_expert_intermediate_size:
            from .shared_experts import SharedExpertMLP

            spec = submodules.shared_experts
            gate = bool(getattr(spec, "params", None) and spec.params.get("gate")) or bool(getattr(config, "moe_shared_expert_gate", False))
            self.shared_experts = SharedExpertMLP(config, spec.submodules if hasattr(spec, "submodules") else spec, gate=gate)

    def set_layer_number(self, n):
        self.layer_number = n
        self.router.set_layer_number(n)

    def forward(self, hidden_states):
        if self.training and ps.get_tensor_model_parallel_world_size() > 1 and not self.config.sequence_parallel:
            raise ValueError("training MoE with tensor parallelism requires sequence parallelism")
        if self.shared_experts is not None:
            self.shared_experts.start(hidden_states)          # side stream under --moe-shared-expert-overlap
        probs, routing_map = self.router(hidden_states)
        dispatched, tokens_per_expert = self.token_dispatcher.token_permutation(hidden_states, probs, routing_map)
        expert_out, mlp_bias = self.experts(dispatched, tokens_per_expert)
        out, mlp_bias = self.token_dispatcher.token_unpermutation(expert_out, mlp_bias)
        if self.shared_experts is not None:
            out = out + self.shared_experts.finish()
        return out, mlp_bias
