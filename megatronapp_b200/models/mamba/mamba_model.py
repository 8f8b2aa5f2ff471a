"""Mamba / hybrid Mamba-attention-MLP language model (reference ``core/models/mamba/mamba_model.py``)."""
from __future__ import annotations

from ...parallel.tp_layers import ColumnParallelLinear, RowParallelLinear
from ...ssm.mamba import MambaStack
from ...transformer.attention import DotProductAttention, SelfAttention, SelfAttentionSubmodules
from ...transformer.mlp import MLP, MLPSubmodules
from ...transformer.norms import FusedLayerNorm, IdentityOp
from ...transformer.spec_utils import ModuleSpec
from ...transformer.transformer_layer import TransformerLayer, TransformerLayerSubmodules
from ..common.embeddings import LanguageModelEmbedding, RotaryEmbedding
from ..common.language_module import LanguageModule


def attention_only_layer_spec():
    return ModuleSpec(module=TransformerLayer, submodules=TransformerLayerSubmodules(
        input_layernorm=FusedLayerNorm,
        self_attention=ModuleSpec(module=SelfAttention, params={"attn_mask_type": "causal"},
                                  submodules=SelfAttentionSubmodules(linear_qkv=ColumnParallelLinear,
                                                                     core_attention=DotProductAttention,
                                                                     linear_proj=RowParallelLinear)),
        pre_mlp_layernorm=IdentityOp, mlp=_ZeroMLP))


def mlp_only_layer_spec():
    return ModuleSpec(module=TransformerLayer, submodules=TransformerLayerSubmodules(
        input_layernorm=IdentityOp, self_attention=_ZeroAttn, pre_mlp_layernorm=FusedLayerNorm,
        mlp=ModuleSpec(module=MLP, submodules=MLPSubmodules(linear_fc1=ColumnParallelLinear, linear_fc2=RowParallelLinear))))


import torch.nn as nn  # noqa: E402


class _ZeroMLP(nn.Module):
    def __init__(self, config=None, **_):
        super().__init__()

    def forward(self, x):
        return x * 0, None


class _ZeroAttn(nn.Module):
    def __init__(self, config=None, **_):
        super().__init__()

    def forward(self, x, **_):
        return x * 0, None


class MambaModel(LanguageModule):
    def __init__(self, config, mamba_stack_spec=None, vocab_size=0, max_sequence_length=0, pre_process=True,
                 hybrid_attention_ratio=0.0, hybrid_mlp_ratio=0.0, hybrid_override_pattern=None, post_process=True,
                 fp16_lm_cross_entropy=False, parallel_output=True, share_embeddings_and_output_weights=False,
                 position_embedding_type="none", rotary_percent=1.0, rotary_base=10000,
                 seq_len_interpolation_factor=None):
        super().__init__(config)
        self.log_constructor_args(locals())
        self.pre_process, self.post_process = pre_process, post_process
        self.share_embeddings_and_output_weights = share_embeddings_and_output_weights
        self.position_embedding_type = position_embedding_type
        if pre_process:
            self.embedding = LanguageModelEmbedding(config, vocab_size, max_sequence_length, position_embedding_type)
        if position_embedding_type == "rope":
            self.rotary_pos_emb = RotaryEmbedding(config.kv_channels, rotary_percent, False, seq_len_interpolation_factor, rotary_base)
        self.decoder = MambaStack(config, attention_only_layer_spec(), mlp_only_layer_spec(), pre_process, post_process,
                                  hybrid_attention_ratio, hybrid_mlp_ratio, hybrid_override_pattern)
        if post_process:
            self.output_layer = ColumnParallelLinear(config.hidden_size, vocab_size, config=config,
                                                     init_method=config.init_method, bias=False, skip_bias_add=False,
                                                     gather_output=not parallel_output,
                                                     skip_weight_param_allocation=pre_process and share_embeddings_and_output_weights)
        if pre_process or post_process:
            self.setup_embeddings_and_output_layer()

    def set_input_tensor(self, t):
        self.decoder.set_input_tensor(t[0] if isinstance(t, (list, tuple)) else t)

    def forward(self, input_ids, position_ids, attention_mask=None, decoder_input=None, labels=None,
                inference_context=None, **_):
        x = decoder_input
        if x is None and self.pre_process:
            x = self.embedding(input_ids=input_ids, position_ids=position_ids)
        rope = None
        if self.position_embedding_type == "rope":
            if inference_context is not None:
                rope = self.rotary_pos_emb(inference_context.max_sequence_length)       # attention slices by offset
            else:
                n = (x if x is not None else self.decoder.input_tensor).size(0)
                rope = self.rotary_pos_emb(n * (self.config.tensor_model_parallel_size if self.config.sequence_parallel else 1))
        h = self.decoder(x, attention_mask=attention_mask, inference_context=inference_context, rotary_pos_emb=rope)
        if not self.post_process:
            return h
        w = self.shared_embedding_or_output_weight() if self.share_embeddings_and_output_weights else None
        logits, _ = self.output_layer(h, weight=w)
        if labels is None:
            return logits.transpose(0, 1).contiguous()
        return self.compute_language_model_loss(labels, logits)
