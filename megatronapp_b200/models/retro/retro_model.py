"""Retro: a GPT decoder whose selected layers cross-attend, chunk by chunk, to encoded retrieved
neighbours (chunked cross-attention), with a small bidirectional encoder that itself cross-attends to
the decoder's chunk representation.

Parity: reference ``core/models/retro/`` -- ``RetroConfig`` (chunk length ``m``, ``k`` neighbours of
retrieved length ``r``, encoder depth, which decoder layers carry cross attention: from layer 6 (or 9)
every 3rd), ``RetroModel.forward(input_ids, position_ids, attention_mask, context_input_ids,
context_position_ids, context_mask, labels)``, ``decoder_attention.py`` (``RetroDecoderCrossAttention``:
the sequence is shifted by ``m - 1`` so a token only sees neighbours of *previous* chunks; the first
retro layer runs the encoder) and ``encoder_attention.py`` (``RetroEncoderCrossAttention``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn.functional as F

from ...config.transformer_config import TransformerConfig
from ...parallel.tp_layers import ColumnParallelLinear, RowParallelLinear
from ...transformer.attention import (CrossAttention, CrossAttentionSubmodules, DotProductAttention, SelfAttention,
                                      SelfAttentionSubmodules)
from ...transformer.mlp import MLP, MLPSubmodules
from ...transformer.module import MegatronModule
from ...transformer.norms import FusedLayerNorm
from ...transformer.spec_utils import ModuleSpec, build_module
from ...transformer.transformer_block import TransformerBlock, TransformerBlockSubmodules
from ...transformer.transformer_layer import TransformerLayer, TransformerLayerSubmodules
from ..gpt.gpt_model import GPTModel


@dataclass
class RetroConfig(TransformerConfig):
    retro_chunk_length: int = 64
    retro_num_neighbors: int = 2
    retro_retrieved_length: int = 128
    retro_encoder_num_layers: int = 2
    retro_encoder_hidden_dropout: float = 0.1
    retro_encoder_attention_dropout: float = 0.1
    retro_num_retrieved_chunks: int = 2
    retro_verify_neighbor_count: bool = True
    retro_attention_gate: float = 1.0      # scales the chunked cross-attention output before the residual add (reference --retro-attention-gate)

    def retro_layer_numbers(self) -> List[int]:
        """1-based decoder layers with chunked cross attention (paper: from 6 (<=15 layers) or 9, every 3rd)."""
        start = 6 if self.num_layers <= 15 else 9
        if self.num_layers < start:
            start = max(1, self.num_layers // 2)
        return list(range(start, self.num_layers + 1, 3))


def _cross_spec(mask_type: str = "no_mask") -> ModuleSpec:
    return ModuleSpec(module=CrossAttention, params={"attn_mask_type": mask_type},
                      submodules=CrossAttentionSubmodules(linear_q=ColumnParallelLinear, linear_kv=ColumnParallelLinear,
                                                          core_attention=DotProductAttention, linear_proj=RowParallelLinear))


def _self_spec(mask_type: str) -> ModuleSpec:
    return ModuleSpec(module=SelfAttention, params={"attn_mask_type": mask_type},
                      submodules=SelfAttentionSubmodules(linear_qkv=ColumnParallelLinear, core_attention=DotProductAttention,
                                                         linear_proj=RowParallelLinear))


def _mlp_spec() -> ModuleSpec:
    return ModuleSpec(module=MLP, submodules=MLPSubmodules(linear_fc1=ColumnParallelLinear, linear_fc2=RowParallelLinear))


class RetroEncoderLayer(TransformerLayer):
    """Bidirectional layer over one neighbour; ``context`` = hidden states of the chunk it was retrieved for."""

    def forward(self, hidden_states, attention_mask=None, context=None, context_mask=None, **kw):
        residual = hidden_states
        x = self.input_layernorm(hidden_states)
        a, ab = self.self_attention(x, attention_mask=attention_mask)
        hidden_states = self._bda(a, ab, residual)
        if self.has_cross and context is not None:
            residual = hidden_states
            x = self.pre_cross_attn_layernorm(hidden_states)
            c, cb = self.cross_attention(x, attention_mask=None, key_value_states=context)
            hidden_states = self._bda(c, cb, residual)
        residual = hidden_states
        m, mb = self.mlp(self.pre_mlp_layernorm(hidden_states))
        return self._bda(m, mb, residual), context


class RetroDecoderLayer(TransformerLayer):
    """GPT layer + chunked cross attention.  ``context`` is a dict carried through the block:
    ``{"neighbor_emb": [r, b*l*k, h] or None, "encoded": [k*r, b*l, h] or None, "encoder": module}``."""

    def __init__(self, config, submodules, layer_number: int = 1, hidden_dropout=None):
        super().__init__(config, submodules, layer_number, hidden_dropout)
        self.m, self.k, self.r = config.retro_chunk_length, config.retro_num_neighbors, config.retro_retrieved_length

    def _chunked_cross_attention(self, hidden_states, ctx):
        s, b, h = hidden_states.shape
        m, k, r = self.m, self.k, self.r
        l = (s + m - 1) // m
        if ctx.get("encoded") is None:
            # first retro layer: encode the neighbours, conditioned on the chunk they belong to
            first_ns = s % m
            chunks = hidden_states if first_ns == 0 else F.pad(hidden_states, (0, 0, 0, 0, 0, l * m - s))
            chunks = chunks.reshape(l, m, b, h).permute(1, 2, 0, 3).reshape(m, b * l, h)             # [m, b*l, h]
            enc_ctx = chunks.unsqueeze(2).expand(m, b * l, k, h).reshape(m, b * l * k, h)            # one copy per neighbour
            enc = ctx["encoder"](ctx["neighbor_emb"], attention_mask=None, context=enc_ctx)          # [r, b*l*k, h]
            ctx["encoded"] = enc.reshape(r, b * l, k, h).permute(2, 0, 1, 3).reshape(k * r, b * l, h)
        # causal alignment: shift left by m - 1 so that the query block that sees chunk i's neighbours is
        # made of the last token of chunk i and the first m - 1 tokens of chunk i + 1
        pad = m - 1
        attending = F.pad(hidden_states[pad:], (0, 0, 0, 0, 0, l * m - (s - pad)))                  # [l*m, b, h]
        attending = attending.reshape(l, m, b, h).permute(1, 2, 0, 3).reshape(m, b * l, h)
        out, bias = self.cross_attention(attending, attention_mask=None, key_value_states=ctx["encoded"])
        if bias is not None:
            out = out + bias
        out = out.reshape(m, b, l, h).permute(2, 0, 1, 3).reshape(l * m, b, h)
        out = F.pad(out, (0, 0, 0, 0, pad, 0))[:s]                                                  # undo the shift
        return out

    def forward(self, hidden_states, attention_mask=None, context=None, context_mask=None, rotary_pos_emb=None, **kw):
        residual = hidden_states
        x = self.input_layernorm(hidden_states)
        a, ab = self.self_attention(x, attention_mask=attention_mask, rotary_pos_emb=rotary_pos_emb)
        hidden_states = self._bda(a, ab, residual)
        if self.has_cross and context is not None and context.get("neighbor_emb") is not None:
            residual = hidden_states
            x = self.pre_cross_attn_layernorm(hidden_states)
            gate = getattr(self.config, "retro_attention_gate", 1.0)
            hidden_states = residual + gate * F.dropout(self._chunked_cross_attention(x, context), self.hidden_dropout, self.training)
        residual = hidden_states
        m_out, mb = self.mlp(self.pre_mlp_layernorm(hidden_states))
        return self._bda(m_out, mb, residual), context


def get_retro_encoder_block_spec(config: RetroConfig) -> TransformerBlockSubmodules:
    layer = ModuleSpec(module=RetroEncoderLayer, submodules=TransformerLayerSubmodules(
        input_layernorm=FusedLayerNorm, self_attention=_self_spec("no_mask"), pre_cross_attn_layernorm=FusedLayerNorm,
        cross_attention=_cross_spec(), pre_mlp_layernorm=FusedLayerNorm, mlp=_mlp_spec()))
    return TransformerBlockSubmodules(layer_specs=[layer] * config.retro_encoder_num_layers, layer_norm=FusedLayerNorm)


def get_retro_decoder_block_spec(config: RetroConfig) -> TransformerBlockSubmodules:
    retro = set(config.retro_layer_numbers())
    specs = []
    for i in range(1, config.num_layers + 1):
        sub = TransformerLayerSubmodules(input_layernorm=FusedLayerNorm, self_attention=_self_spec("causal"),
                                         pre_mlp_layernorm=FusedLayerNorm, mlp=_mlp_spec())
        if i in retro:
            sub.pre_cross_attn_layernorm = FusedLayerNorm
            sub.cross_attention = _cross_spec()
        specs.append(ModuleSpec(module=RetroDecoderLayer, submodules=sub))
    return TransformerBlockSubmodules(layer_specs=specs, layer_norm=FusedLayerNorm)


class RetroModel(GPTModel):
    def __init__(self, config: RetroConfig, transformer_layer_spec=None, vocab_size: int = 0, max_sequence_length: int = 0,
                 **kw):
        spec = transformer_layer_spec or get_retro_decoder_block_spec(config)
        super().__init__(config, spec, vocab_size, max_sequence_length, **kw)
        import copy

        enc_cfg = copy.copy(config)
        enc_cfg.num_layers = config.retro_encoder_num_layers
        enc_cfg.hidden_dropout = config.retro_encoder_hidden_dropout
        enc_cfg.attention_dropout = config.retro_encoder_attention_dropout
        enc_cfg.pipeline_model_parallel_size = 1
        self.retro_encoder = TransformerBlock(enc_cfg, get_retro_encoder_block_spec(config), pre_process=True, post_process=True)

    def forward(self, input_ids, position_ids, attention_mask=None, context_input_ids=None, context_position_ids=None,
                context_mask=None, decoder_input=None, labels=None, inference_context=None, **kw):
        """``context_input_ids``: ``[b * l * k, r]`` (or ``[b, l, k, r]``) retrieved neighbour tokens."""
        ctx = None
        if context_input_ids is not None:
            ci = context_input_ids.reshape(-1, context_input_ids.shape[-1])
            cp = context_position_ids.reshape(-1, ci.shape[-1]) if context_position_ids is not None else \
                torch.arange(ci.shape[-1], device=ci.device).unsqueeze(0).expand_as(ci)
            ctx = {"neighbor_emb": self.embedding(input_ids=ci, position_ids=cp), "encoded": None, "encoder": self.retro_encoder}
        if decoder_input is None and self.pre_process:
            decoder_input = self.embedding(input_ids=input_ids, position_ids=position_ids)
        rotary = None
        if self.position_embedding_type == "rope":
            rotary = self.rotary_pos_emb(decoder_input.shape[0])
        hidden = self.decoder(hidden_states=decoder_input, attention_mask=attention_mask, context=ctx, rotary_pos_emb=rotary)
        if not self.post_process:
            return hidden
        w = self.shared_embedding_or_output_weight() if self.share_embeddings_and_output_weights else None
        logits, _ = self.output_layer(hidden, weight=w)
        if labels is None:
            return logits.transpose(0, 1).contiguous()
        return self.compute_language_model_loss(labels, logits)
