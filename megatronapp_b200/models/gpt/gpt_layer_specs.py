"""Layer specs for GPT-family models (dense, MoE, heterogeneous patterns).

Parity: reference ``models/gpt/gpt_layer_specs.py`` (``get_gpt_layer_local_spec``,
``get_gpt_layer_with_transformer_engine_spec`` -- here both resolve to the same B200-native
layer, ``get_gpt_decoder_block_spec`` for ``moe_layer_freq`` patterns, ``get_mlp_module_spec``).
"""
from __future__ import annotations

from typing import Optional

from ...parallel.tp_layers import ColumnParallelLinear, RowParallelLinear
from ...transformer.attention import DotProductAttention, SelfAttention, SelfAttentionSubmodules
from ...transformer.mlp import MLP, MLPSubmodules
from ...transformer.norms import FusedLayerNorm, IdentityOp
from ...transformer.spec_utils import ModuleSpec
from ...transformer.transformer_block import TransformerBlockSubmodules
from ...transformer.transformer_layer import TransformerLayer, TransformerLayerSubmodules, get_num_layers_to_build, get_transformer_layer_offset


def get_mlp_module_spec(use_te: bool = False, num_experts: Optional[int] = None, moe_grouped_gemm: bool = False,
                        fp8=None, moe_use_legacy_grouped_gemm: bool = False) -> ModuleSpec:
    if num_experts is None:
        return ModuleSpec(module=MLP, submodules=MLPSubmodules(linear_fc1=ColumnParallelLinear,
                                                               linear_fc2=RowParallelLinear))
    from ...moe.moe_layer import MoELayer, MoESubmodules
    from ...moe.experts import GroupedMLP, SequentialMLP

    experts = ModuleSpec(module=GroupedMLP if moe_grouped_gemm else SequentialMLP,
                         submodules=MLPSubmodules(linear_fc1=ColumnParallelLinear, linear_fc2=RowParallelLinear))
    from ...moe.shared_experts import SharedExpertMLP

    shared = ModuleSpec(module=SharedExpertMLP, params={"gate": False},
                        submodules=MLPSubmodules(linear_fc1=ColumnParallelLinear, linear_fc2=RowParallelLinear))
    return ModuleSpec(module=MoELayer, submodules=MoESubmodules(experts=experts, shared_experts=shared))


def get_gpt_layer_local_spec(num_experts: Optional[int] = None, moe_grouped_gemm: bool = False,
                             qk_layernorm: bool = False, multi_latent_attention: bool = False,
                             fp8=None, moe_use_legacy_grouped_gemm: bool = False, normalization=None) -> ModuleSpec:
    mlp = get_mlp_module_spec(False, num_experts, moe_grouped_gemm)
    if multi_latent_attention:
        from ...transformer.mla import MLASelfAttention, MLASelfAttentionSubmodules

        attn = ModuleSpec(module=MLASelfAttention, params={"attn_mask_type": "causal"},
                          submodules=MLASelfAttentionSubmodules(
                              linear_q_proj=ColumnParallelLinear, linear_q_down_proj=ColumnParallelLinear,
                              linear_q_up_proj=ColumnParallelLinear, linear_kv_down_proj=ColumnParallelLinear,
                              linear_kv_up_proj=ColumnParallelLinear, core_attention=DotProductAttention,
                              linear_proj=RowParallelLinear,
                              q_layernorm=FusedLayerNorm if qk_layernorm else IdentityOp,
                              kv_layernorm=FusedLayerNorm if qk_layernorm else IdentityOp))
    else:
        attn = ModuleSpec(module=SelfAttention, params={"attn_mask_type": "causal"},
                          submodules=SelfAttentionSubmodules(
                              linear_qkv=ColumnParallelLinear, core_attention=DotProductAttention,
                              linear_proj=RowParallelLinear,
                              q_layernorm=FusedLayerNorm if qk_layernorm else None,
                              k_layernorm=FusedLayerNorm if qk_layernorm else None))
    return ModuleSpec(module=TransformerLayer, submodules=TransformerLayerSubmodules(
        input_layernorm=FusedLayerNorm, self_attention=attn, pre_mlp_layernorm=FusedLayerNorm, mlp=mlp))


def get_gpt_layer_with_transformer_engine_spec(num_experts=None, moe_grouped_gemm=False, qk_layernorm=False,
                                               multi_latent_attention=False, fp8=None,
                                               moe_use_legacy_grouped_gemm=False) -> ModuleSpec:
    """``--transformer-impl transformer_engine`` resolves to the same native layer: there is no
    separate TE path in this framework (fused attention, fp8 and overlap live in our kernels)."""
    return get_gpt_layer_local_spec(num_experts, moe_grouped_gemm, qk_layernorm, multi_latent_attention, fp8)


def get_gpt_decoder_block_spec(config, use_transformer_engine: bool = False) -> TransformerBlockSubmodules:
    dense = get_gpt_layer_local_spec(None, False, config.qk_layernorm, config.multi_latent_attention)
    moe = get_gpt_layer_local_spec(config.num_moe_experts, config.moe_grouped_gemm, config.qk_layernorm,
                                   config.multi_latent_attention) if config.num_moe_experts else dense
    freq = config.moe_layer_freq
    if isinstance(freq, int):
        pattern = [1 if (i % freq == 0) else 0 for i in range(config.num_layers)]
    else:
        pattern = list(freq)
    if config.num_moe_experts is None:
        pattern = [0] * config.num_layers
    specs = [moe if p else dense for p in pattern]
    off = get_transformer_layer_offset(config)
    n = get_num_layers_to_build(config)
    return TransformerBlockSubmodules(layer_specs=specs[off:off + n], layer_norm=FusedLayerNorm)


def get_moe_module_spec(use_te: bool = True, num_experts: Optional[int] = None, moe_grouped_gemm: bool = False,
                        moe_use_legacy_grouped_gemm: bool = False) -> ModuleSpec:
    """``MoELayer`` spec with its experts (reference ``moe_module_specs.get_moe_module_spec``)."""
    if num_experts is None:
        raise ValueError("num_experts is required for an MoE module spec")
    return get_mlp_module_spec(use_te=False, num_experts=num_experts, moe_grouped_gemm=moe_grouped_gemm)


def get_gpt_mtp_block_spec(config, spec, use_transformer_engine: bool = False):
    """Multi-token-prediction block spec: ``config.mtp_num_layers`` copies of the decoder layer spec (``None`` without MTP)."""
    from ...transformer.multi_token_prediction import get_mtp_block_spec

    return get_mtp_block_spec(config, spec)
