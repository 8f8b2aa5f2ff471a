"""Heterogeneous (NAS-style) GPT blocks: every layer may change its attention / MLP shape or drop it.

Parity: reference ``core/models/gpt/heterogeneous/heterogeneous_layer_specs.py`` and
``transformer/heterogeneous/heterogeneous_config.py`` (``--heterogeneous-layers-config-path`` JSON in
the Nemotron/Puzzle "block_configs" format: per layer ``attention: {no_op, replace_with_linear,
n_heads_in_group | num_query_groups}``, ``ffn: {no_op, replace_with_linear, ffn_mult |
intermediate_size}``).
"""
from __future__ import annotations

import copy
import json
from dataclasses import dataclass
from typing import List, Optional

import torch

from ....parallel.tp_layers import ColumnParallelLinear
from ....transformer.module import MegatronModule
from ....transformer.norms import FusedLayerNorm, IdentityOp
from ....transformer.spec_utils import ModuleSpec
from ....transformer.transformer_block import TransformerBlockSubmodules
from ....transformer.transformer_layer import TransformerLayer, get_num_layers_to_build, get_transformer_layer_offset
from ..gpt_layer_specs import get_gpt_layer_local_spec


@dataclass
class BlockConfig:
    attn_no_op: bool = False
    attn_linear: bool = False
    num_query_groups: Optional[int] = None
    mlp_no_op: bool = False
    mlp_linear: bool = False
    ffn_hidden_size: Optional[int] = None


def _ffn_mult_to_size(mult: float, hidden: int, multiple_of: int = 256) -> int:
    size = int(2 * mult * hidden / 3)
    return (size + multiple_of - 1) // multiple_of * multiple_of


def parse_heterogeneous_config(cfg, num_attention_heads: int, hidden_size: int) -> List[BlockConfig]:
    """``cfg``: path to / dict of the JSON with a ``block_configs`` list."""
    if isinstance(cfg, str):
        with open(cfg) as f:
            cfg = json.load(f)
    out = []
    for b in cfg["block_configs"]:
        a, m = b.get("attention", {}) or {}, b.get("ffn", b.get("mlp", {})) or {}
        ng = a.get("num_query_groups")
        if ng is None and a.get("n_heads_in_group"):
            ng = num_attention_heads // int(a["n_heads_in_group"])
        ffn = m.get("intermediate_size") or m.get("ffn_hidden_size")
        if ffn is None and m.get("ffn_mult") is not None:
            ffn = _ffn_mult_to_size(float(m["ffn_mult"]), hidden_size)
        out.append(BlockConfig(bool(a.get("no_op")), bool(a.get("replace_with_linear")), ng, bool(m.get("no_op")),
                               bool(m.get("replace_with_linear")), ffn))
    return out


class _LinearBranch(MegatronModule):
    """``replace_with_linear``: the sub-block collapses to one hidden -> hidden projection."""

    def __init__(self, config, **_):
        super().__init__(config)
        self.linear = ColumnParallelLinear(config.hidden_size, config.hidden_size, config=config, init_method=config.init_method,
                                           bias=False, gather_output=True, skip_bias_add=True)

    def forward(self, hidden_states, *a, **k):
        return self.linear(hidden_states)


class HeterogeneousLayer(TransformerLayer):
    """A ``TransformerLayer`` built from a per-layer copy of the config."""

    def __init__(self, config, submodules, layer_number: int = 1, hidden_dropout=None, block: Optional[BlockConfig] = None):
        cfg = copy.copy(config)
        block = block or BlockConfig()
        if block.num_query_groups:
            cfg.num_query_groups = block.num_query_groups
        if block.ffn_hidden_size:
            cfg.ffn_hidden_size = block.ffn_hidden_size
        sub = copy.copy(submodules)
        if block.attn_no_op:
            sub.input_layernorm, sub.self_attention = IdentityOp, IdentityOp
        elif block.attn_linear:
            sub.self_attention = ModuleSpec(module=_LinearBranch)
        if block.mlp_no_op:
            sub.pre_mlp_layernorm, sub.mlp = IdentityOp, IdentityOp
        elif block.mlp_linear:
            sub.mlp = ModuleSpec(module=_LinearBranch)
        self.block = block
        super().__init__(cfg, sub, layer_number, hidden_dropout)

    def forward(self, hidden_states, attention_mask=None, rotary_pos_emb=None, inference_context=None, **kw):
        b = self.block
        if not b.attn_no_op:
            residual = hidden_states
            x = self.input_layernorm(hidden_states)
            if b.attn_linear:
                a, ab = self.self_attention(x)
            else:
                a, ab = self.self_attention(x, attention_mask=attention_mask, rotary_pos_emb=rotary_pos_emb,
                                            inference_context=inference_context)
            hidden_states = self._bda(a, ab, residual)
        if not b.mlp_no_op:
            residual = hidden_states
            m, mb = self.mlp(self.pre_mlp_layernorm(hidden_states))
            hidden_states = self._bda(m, mb, residual)
        return hidden_states, None


def get_gpt_heterogeneous_layer_spec(config, blocks: List[BlockConfig]) -> TransformerBlockSubmodules:
    assert len(blocks) == config.num_layers, "one block config per layer"
    base = get_gpt_layer_local_spec(None, False, config.qk_layernorm)
    off, n = get_transformer_layer_offset(config), get_num_layers_to_build(config)
    specs = [ModuleSpec(module=HeterogeneousLayer, params={"block": blocks[i]}, submodules=base.submodules)
             for i in range(off, off + n)]
    return TransformerBlockSubmodules(layer_specs=specs, layer_norm=FusedLayerNorm)


# -- configuration classes of the reference's ``transformer/heterogeneous/heterogeneous_config.py`` ------------------------------------
@dataclass
class AttentionConfig:
    """One block's attention: absent (``no_op``), replaced by a linear map, or grouped-query attention with its own group count."""
    no_op: bool = False
    replace_with_linear: bool = False
    num_query_groups: Optional[int] = None

    @classmethod
    def build_config_from_dict(cls, d: dict, num_attention_heads: int) -> "AttentionConfig":
        ng = d.get("num_query_groups")
        if ng is None and d.get("n_heads_in_group"):
            ng = num_attention_heads // int(d["n_heads_in_group"])
        return cls(bool(d.get("no_op")), bool(d.get("replace_with_linear")), ng)


@dataclass
class MLPConfig:
    no_op: bool = False
    replace_with_linear: bool = False
    ffn_hidden_size: Optional[int] = None

    @classmethod
    def build_config_from_dict(cls, d: dict, hidden_size: int) -> "MLPConfig":
        ffn = d.get("intermediate_size") or d.get("ffn_hidden_size")
        if ffn is None and d.get("ffn_mult") is not None:
            ffn = _ffn_mult_to_size(float(d["ffn_mult"]), hidden_size)
        return cls(bool(d.get("no_op")), bool(d.get("replace_with_linear")), ffn)


@dataclass
class TransformerBlockConfig:
    attention: AttentionConfig
    mlp: MLPConfig

    @classmethod
    def build_from_dict(cls, block: dict, num_attention_heads: int, hidden_size: int) -> "TransformerBlockConfig":
        return cls(AttentionConfig.build_config_from_dict(block.get("attention", {}) or {}, num_attention_heads),
                   MLPConfig.build_config_from_dict(block.get("ffn", block.get("mlp", {})) or {}, hidden_size))

    def to_block_config(self) -> BlockConfig:
        a, m = self.attention, self.mlp
        return BlockConfig(a.no_op, a.replace_with_linear, a.num_query_groups, m.no_op, m.replace_with_linear, m.ffn_hidden_size)


def HeterogeneousTransformerConfig(heterogeneous_layers_config_path=None, heterogeneous_layers_config_encoded_json=None, **kwargs):
    """A ``TransformerConfig`` with the per-block description attached: ``per_block_parameters`` (list of
    ``TransformerBlockConfig``) parsed from the JSON file / string, plus the two source attributes the reference's class has."""
    from ....config.transformer_config import TransformerConfig

    cfg = TransformerConfig(**kwargs)
    cfg.heterogeneous_layers_config_path = heterogeneous_layers_config_path
    cfg.heterogeneous_layers_config_encoded_json = heterogeneous_layers_config_encoded_json
    data = None
    if heterogeneous_layers_config_encoded_json:
        data = json.loads(heterogeneous_layers_config_encoded_json)
    elif heterogeneous_layers_config_path:
        with open(heterogeneous_layers_config_path) as f:
            data = json.load(f)
    cfg.per_block_parameters = [] if data is None else [
        TransformerBlockConfig.build_from_dict(b, cfg.num_attention_heads, cfg.hidden_size) for b in data["block_configs"]]
    return cfg
