"""Stack of transformer layers owned by one (virtual) pipeline stage.

Parity: reference ``transformer/transformer_block.py`` (``TransformerBlock`` :220: builds
``num_layers/pp(/vpp)`` layers, full activation recompute ``uniform``/``block``, final norm on
the post-process stage, ``set_input_tensor`` for pipelining, MegaScope system perturbation
:542-544).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Union

import torch
import torch.nn as nn

from ..parallel import state as ps
from ..parallel.random import checkpoint as recompute
from ..scan.tracer import scope as _scope
from ..scope import hooks as scope_hooks
from .module import MegatronModule
from .norms import FusedLayerNorm
from .spec_utils import ModuleSpec, build_module
from .transformer_layer import get_num_layers_to_build


@dataclass
class TransformerBlockSubmodules:
    layer_specs: List[ModuleSpec] = None
    layer_norm: Optional[Union[ModuleSpec, type]] = None


def _get_block_submodules(config, spec) -> TransformerBlockSubmodules:
    if isinstance(spec, TransformerBlockSubmodules):
        return spec
    if isinstance(spec, ModuleSpec):
        from .transformer_layer import TransformerLayer

        if spec.module is TransformerBlock:
            return spec.submodules
        n = get_num_layers_to_build(config)
        return TransformerBlockSubmodules(layer_specs=[spec] * n, layer_norm=FusedLayerNorm)
    raise TypeError(f"unsupported block spec {type(spec)}")


class TransformerBlock(MegatronModule):
    def __init__(self, config, spec, post_layer_norm: bool = True, pre_process: bool = True,
                 post_process: bool = True):
        super().__init__(config)
        self.submodules = _get_block_submodules(config, spec)
        self.post_layer_norm, self.pre_process, self.post_process = post_layer_norm, pre_process, post_process
        self.input_tensor = None
        self.layers = nn.ModuleList([build_module(s, config=config, layer_number=i + 1)
                                     for i, s in enumerate(self.submodules.layer_specs)])
        if self.submodules.layer_norm and self.post_process and self.post_layer_norm:
            self.final_layernorm = build_module(self.submodules.layer_norm, config=config,
                                                hidden_size=config.hidden_size, eps=config.layernorm_epsilon)
        else:
            self.final_layernorm = None

    def set_input_tensor(self, input_tensor):
        self.input_tensor = input_tensor

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        """Checkpoint keys use the GLOBAL layer index (``layer_number - 1`` already includes the pipeline / virtual
        pipeline offset), so a checkpoint written at one PP degree loads at another."""
        from .module import sharded_state_dict_of

        sd = {}
        for i, layer in enumerate(self.layers):
            g = getattr(layer, "layer_number", i + 1) - 1
            sd.update(sharded_state_dict_of(layer, f"{prefix}layers.{g}.", sharded_offsets, metadata))
        if self.final_layernorm is not None:
            sd.update(sharded_state_dict_of(self.final_layernorm, f"{prefix}final_layernorm.", sharded_offsets, metadata))
        return sd

    def _checkpointed_forward(self, hidden_states, attention_mask, context, context_mask, rotary_pos_emb,
                              packed_seq_params):
        cfg = self.config

        def custom(start, end):
            def fwd(h, m, c, cm, rope):
                for i in range(start, end):
                    h, c = self.layers[i](h, attention_mask=m, context=c, context_mask=cm, rotary_pos_emb=rope,
                                          packed_seq_params=packed_seq_params)
                return h
            return fwd

        n = len(self.layers)
        if cfg.recompute_method == "uniform":
            i = 0
            while i < n:
                j = min(n, i + cfg.recompute_num_layers)
                hidden_states = recompute(custom(i, j), cfg.distribute_saved_activations, hidden_states,
                                          attention_mask, context, context_mask, rotary_pos_emb)
                i = j
        elif cfg.recompute_method == "block":
            for i in range(n):
                if i < cfg.recompute_num_layers:
                    hidden_states = recompute(custom(i, i + 1), cfg.distribute_saved_activations, hidden_states,
                                              attention_mask, context, context_mask, rotary_pos_emb)
                else:
                    hidden_states = custom(i, i + 1)(hidden_states, attention_mask, context, context_mask,
                                                     rotary_pos_emb)
        else:
            raise ValueError("invalid activation recompute method")
        return hidden_states

    def forward(self, hidden_states, attention_mask=None, context=None, context_mask=None, rotary_pos_emb=None,
                rotary_pos_cos=None, rotary_pos_sin=None, attention_bias=None, inference_context=None,
                packed_seq_params=None, sequence_len_offset=None, inference_params=None):
        inference_context = inference_context or inference_params
        if not self.pre_process:
            hidden_states = self.input_tensor
        if self.config.recompute_granularity == "full" and self.training:
            hidden_states = self._checkpointed_forward(hidden_states, attention_mask, context, context_mask,
                                                       rotary_pos_emb, packed_seq_params)
        else:
            graphable = (self.config.enable_cuda_graph and inference_context is None and attention_mask is None
                         and context is None and packed_seq_params is None and hidden_states.is_cuda)
            if graphable and not hasattr(self.layers[0], "cuda_graph_manager"):
                from .cuda_graphs import graph_transformer_block

                graph_transformer_block(self, self.config.cuda_graph_warmup_steps)
            if graphable:
                from .cuda_graphs import capture_block_upfront

                extra = (rotary_pos_emb,) if torch.is_tensor(rotary_pos_emb) else ()
                capture_block_upfront(self, hidden_states, *extra)
            for layer in self.layers:
                if graphable:
                    hidden_states = layer.cuda_graph_manager(hidden_states, *extra)
                    hidden_states = scope_hooks.perturb_system(hidden_states)
                    continue
                hidden_states, context = layer(
                    hidden_states, attention_mask=attention_mask, context=context, context_mask=context_mask,
                    rotary_pos_emb=rotary_pos_emb, attention_bias=attention_bias,
                    inference_context=inference_context, packed_seq_params=packed_seq_params)
                hidden_states = scope_hooks.perturb_system(hidden_states)
        if self.final_layernorm is not None:
            hidden_states = self.final_layernorm(hidden_states)
        return hidden_states
