"""Decoder layer specs for the LLaVA language model (reference ``core/models/multimodal/llava_spec.py``): the GPT layer with
a causal mask, optionally with query / key LayerNorm (``qk_layernorm``) as used by the larger NVLM decoders."""
from typing import Optional

from ...transformer.spec_utils import ModuleSpec
from ..gpt.gpt_layer_specs import get_gpt_layer_local_spec


def decoder_model_with_local_default_spec(num_experts: Optional[int] = None, moe_grouped_gemm: bool = False,
                                          qk_layernorm: bool = False) -> ModuleSpec:
    return get_gpt_layer_local_spec(num_experts, moe_grouped_gemm, qk_layernorm=qk_layernorm)


def decoder_model_with_transformer_engine_default_spec(num_experts: Optional[int] = None, moe_grouped_gemm: bool = False,
                                                       qk_layernorm: bool = False) -> ModuleSpec:
    return decoder_model_with_local_default_spec(num_experts, moe_grouped_gemm, qk_layernorm)
