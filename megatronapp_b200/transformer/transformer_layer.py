"""Pre-LN transformer layer and the layer-offset arithmetic.

Parity: reference ``transformer/transformer_layer.py`` (``TransformerLayer`` :237, forward
:382-592: input_layernorm -> self_attention -> bias-dropout-add -> [cross attention] ->
pre_mlp_layernorm -> mlp -> bias-dropout-add; ``get_transformer_layer_offset``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
import dataclasses
from typing import Dict, Optional, Union

import torch

from ..parallel import state as ps
from ..parallel.topology import layer_offset
from .. import ops
from ..scan.tracer import scope as _scope
from ..scope import hooks as scope_hooks
from .module import MegatronModule
from .norms import IdentityOp
from .spec_utils import ModuleSpec, build_module


def get_transformer_layer_offset(config) -> int:
    pp = ps.get_pipeline_model_parallel_world_size()
    if pp == 1:
        return 0
    vpp = config.virtual_pipeline_model_parallel_size
    off, _ = layer_offset(config.num_layers, pp, ps.get_pipeline_model_parallel_rank(), vpp,
                          ps.get_virtual_pipeline_model_parallel_rank() if vpp else None,
                          config.num_layers_in_first_pipeline_stage, config.num_layers_in_last_pipeline_stage)
    return off


def get_num_layers_to_build(config) -> int:
    pp = ps.get_pipeline_model_parallel_world_size()
    vpp = config.virtual_pipeline_model_parallel_size
    _, cnt = layer_offset(config.num_layers, pp, ps.get_pipeline_model_parallel_rank(), vpp,
                          ps.get_virtual_pipeline_model_parallel_rank() if vpp else None,
                          config.num_layers_in_first_pipeline_stage, config.num_layers_in_last_pipeline_stage)
    return cnt


@dataclass
class TransformerLayerSubmodules:
    input_layernorm: Union[ModuleSpec, type] = IdentityOp
    self_attention: Union[ModuleSpec, type] = IdentityOp
    self_attn_bda: Union[ModuleSpec, type] = None
    pre_cross_attn_layernorm: Union[ModuleSpec, type] = IdentityOp
    cross_attention: Union[ModuleSpec, type] = IdentityOp
    cross_attn_bda: Union[ModuleSpec, type] = None
    pre_mlp_layernorm: Union[ModuleSpec, type] = IdentityOp
    mlp: Union[ModuleSpec, type] = IdentityOp
    mlp_bda: Union[ModuleSpec, type] = None
    sharded_state_dict_keys_map: Dict[str, str] = field(default_factory=dict)


def layer_precision_config(config, layer_number: int):
    """--first-last-layers-bf16 (reference ``fp8_utils.is_first_last_bf16_layer``): the first
    ``num_layers_at_start_in_bf16`` and the last ``num_layers_at_end_in_bf16`` layers of an fp8 model keep bf16 GEMMs."""
    if not (getattr(config, "fp8", None) and getattr(config, "first_last_layers_bf16", False)):
        return config
    n = config.num_layers
    if layer_number <= config.num_layers_at_start_in_bf16 or layer_number > n - config.num_layers_at_end_in_bf16:
        return dataclasses.replace(config, fp8=None)
    return config


class TransformerLayer(MegatronModule):
    def __init__(self, config, submodules: TransformerLayerSubmodules, layer_number: int = 1,
                 hidden_dropout: Optional[float] = None):
        self.layer_number = layer_number + get_transformer_layer_offset(config)
        config = layer_precision_config(config, self.layer_number)
        super().__init__(config)
        self.hidden_dropout = config.hidden_dropout if hidden_dropout is None else hidden_dropout
        eps = config.layernorm_epsilon
        self.input_layernorm = build_module(submodules.input_layernorm, config=config,
                                            hidden_size=config.hidden_size, eps=eps)
        self.self_attention = build_module(submodules.self_attention, config=config, layer_number=self.layer_number)
        self.pre_cross_attn_layernorm = build_module(submodules.pre_cross_attn_layernorm, config=config,
                                                     hidden_size=config.hidden_size, eps=eps)
        self.cross_attention = build_module(submodules.cross_attention, config=config, layer_number=self.layer_number)
        self.has_cross = not isinstance(self.cross_attention, IdentityOp)
        self.pre_mlp_layernorm = build_module(submodules.pre_mlp_layernorm, config=config,
                                              hidden_size=config.hidden_size, eps=eps)
        self.mlp = build_module(submodules.mlp, config=config)
        if hasattr(self.mlp, "set_layer_number"):
            self.mlp.set_layer_number(self.layer_number)
        if hasattr(self.mlp, "layer_number"):
            self.mlp.layer_number = self.layer_number
        self.recompute_mlp = (config.recompute_granularity == "selective"
                              and "mlp" in (config.recompute_modules or []))

    def _bda(self, x, bias, residual):
        return ops.bias_dropout_add(x, bias, residual, self.hidden_dropout, self.training)

    def forward(self, hidden_states, attention_mask=None, context=None, context_mask=None,
                rotary_pos_emb=None, rotary_pos_cos=None, rotary_pos_sin=None, attention_bias=None,
                inference_context=None, packed_seq_params=None, sequence_len_offset=None,
                inference_params=None):
        inference_context = inference_context or inference_params
        with _scope("transformer_layer", layer=self.layer_number):
            hidden_states = ops.slow_spin(hidden_states)      # MegaScan slow-rank injection (no-op unless armed)
            with _scope("_forward_attention"):
                residual = hidden_states
                x = self.input_layernorm(hidden_states)
                attn_out, attn_bias = self.self_attention(
                    x, attention_mask=attention_mask, inference_context=inference_context,
                    rotary_pos_emb=rotary_pos_emb, attention_bias=attention_bias,
                    packed_seq_params=packed_seq_params)
                hidden_states = self._bda(attn_out, attn_bias, residual)
                if self.has_cross:
                    residual = hidden_states
                    x = self.pre_cross_attn_layernorm(hidden_states)
                    ca_out, ca_bias = self.cross_attention(x, attention_mask=context_mask,
                                                           key_value_states=context,
                                                           inference_context=inference_context)
                    hidden_states = self._bda(ca_out, ca_bias, residual)
            hidden_states = ops.slow_spin(hidden_states)      # second injection point: a slow GPU is late at every collective
            with _scope("_forward_mlp"):
                residual = hidden_states
                x = self.pre_mlp_layernorm(hidden_states)
                if self.recompute_mlp and self.training:
                    from ..parallel.random import checkpoint

                    with _scope("recompute_mlp"):
                        mlp_out = checkpoint(lambda t: self.mlp(t)[0], False, x)
                        mlp_bias = getattr(getattr(self.mlp, "linear_fc2", None), "bias", None)
                else:
                    with _scope("mlp"):
                        mlp_out, mlp_bias = self.mlp(x)
                mlp_out = scope_hooks.perturb_calculation(mlp_out)
                hidden_states = self._bda(mlp_out, mlp_bias, residual)
        return hidden_states, context

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        return super().sharded_state_dict(prefix, sharded_offsets, metadata)


class BaseTransformerLayer:
    """Marker base class (reference ``BaseTransformerLayer``): layers implementing the ``TransformerLayer`` call contract --
    ``forward(hidden_states, attention_mask, context, context_mask, rotary_pos_emb, inference_context, packed_seq_params)``
    returning ``(hidden_states, context)`` -- can be placed in a ``TransformerBlock`` spec."""


class MLPLayer(TransformerLayer):
    """A transformer layer reduced to its MLP half (pre-norm -> MLP -> residual): the ``-`` symbol of hybrid Mamba stacks
    (reference ``ssm/mlp_layer.py``).  Built from a spec whose attention submodules are ``IdentityOp``."""

    def forward(self, hidden_states, attention_mask=None, **kwargs):
        residual = hidden_states
        out, bias = self.mlp(self.pre_mlp_layernorm(hidden_states))
        return self._bda(out, bias, residual), kwargs.get("context")
