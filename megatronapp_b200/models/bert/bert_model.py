"""BERT (bidirectional encoder + MLM head + optional NSP binary head).

Parity: reference ``core/models/bert/`` (``BertModel``, ``BertLMHead``, ``Pooler``;
``bert_extended_attention_mask`` padding masks; tokentype embeddings; tied output weights)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ...parallel.tp_layers import ColumnParallelLinear, RowParallelLinear
from ...transformer.attention import DotProductAttention, SelfAttention, SelfAttentionSubmodules
from ...transformer.mlp import MLP, MLPSubmodules
from ...transformer.norms import FusedLayerNorm
from ...transformer.spec_utils import ModuleSpec
from ...transformer.transformer_block import TransformerBlock
from ...transformer.transformer_layer import TransformerLayer, TransformerLayerSubmodules
from ... import ops
from ..common.embeddings import LanguageModelEmbedding
from ..common.language_module import LanguageModule


def bert_layer_local_spec() -> ModuleSpec:
    return ModuleSpec(module=TransformerLayer, submodules=TransformerLayerSubmodules(
        input_layernorm=FusedLayerNorm,
        self_attention=ModuleSpec(module=SelfAttention, params={"attn_mask_type": "padding"},
                                  submodules=SelfAttentionSubmodules(linear_qkv=ColumnParallelLinear,
                                                                     core_attention=DotProductAttention,
                                                                     linear_proj=RowParallelLinear)),
        pre_mlp_layernorm=FusedLayerNorm,
        mlp=ModuleSpec(module=MLP, submodules=MLPSubmodules(linear_fc1=ColumnParallelLinear, linear_fc2=RowParallelLinear))))


def bert_extended_attention_mask(attention_mask: torch.Tensor) -> torch.Tensor:
    """[b, s] (1 = real token) -> [b, 1, s, s] bool, True = masked out."""
    m = attention_mask.unsqueeze(1) * attention_mask.unsqueeze(2)
    return (m < 0.5).unsqueeze(1)


def bert_position_ids(token_ids):
    s = token_ids.size(1)
    return torch.arange(s, dtype=torch.long, device=token_ids.device).unsqueeze(0).expand_as(token_ids)


class Pooler(nn.Module):
    def __init__(self, hidden_size, init_method, config, sequence_parallel=False):
        super().__init__()
        self.dense = nn.Linear(hidden_size, hidden_size, dtype=config.params_dtype)
        init_method(self.dense.weight)
        nn.init.zeros_(self.dense.bias)
        self.sequence_parallel = sequence_parallel

    def forward(self, hidden_states, sequence_index=0):
        if self.sequence_parallel:
            from ...parallel.mappings import gather_from_sequence_parallel_region

            hidden_states = gather_from_sequence_parallel_region(hidden_states, tensor_parallel_output_grad=False)
        return torch.tanh(self.dense(hidden_states[sequence_index]))


class BertLMHead(nn.Module):
    def __init__(self, hidden_size, config):
        super().__init__()
        self.dense = nn.Linear(hidden_size, hidden_size, dtype=config.params_dtype)
        config.init_method(self.dense.weight)
        nn.init.zeros_(self.dense.bias)
        self.layer_norm = FusedLayerNorm(config, hidden_size, config.layernorm_epsilon)

    def forward(self, hidden_states):
        return self.layer_norm(ops.bias_gelu(self.dense(hidden_states)))


class BertModel(LanguageModule):
    def __init__(self, config, num_tokentypes: int, transformer_layer_spec, vocab_size: int, max_sequence_length: int,
                 pre_process=True, post_process=True, fp16_lm_cross_entropy=False, parallel_output=True,
                 share_embeddings_and_output_weights=True, position_embedding_type="learned_absolute",
                 add_binary_head=True, return_embeddings=False):
        super().__init__(config)
        self.log_constructor_args(locals())
        self.pre_process, self.post_process = pre_process, post_process
        self.add_binary_head, self.return_embeddings = add_binary_head, return_embeddings
        self.share_embeddings_and_output_weights = share_embeddings_and_output_weights
        self.parallel_output = parallel_output
        self.vocab_size, self.max_sequence_length = vocab_size, max_sequence_length
        if pre_process:
            self.embedding = LanguageModelEmbedding(config, vocab_size, max_sequence_length, position_embedding_type,
                                                    num_tokentypes=num_tokentypes)
        self.encoder = TransformerBlock(config, transformer_layer_spec, pre_process=pre_process, post_process=post_process)
        if post_process:
            self.lm_head = BertLMHead(config.hidden_size, config)
            self.output_layer = ColumnParallelLinear(config.hidden_size, vocab_size, config=config,
                                                     init_method=config.init_method, bias=True, skip_bias_add=False,
                                                     gather_output=not parallel_output,
                                                     skip_weight_param_allocation=pre_process and share_embeddings_and_output_weights)
            if add_binary_head:
                self.pooler = Pooler(config.hidden_size, config.init_method, config, config.sequence_parallel)
                self.binary_head = nn.Linear(config.hidden_size, 2, dtype=config.params_dtype)
                config.init_method(self.binary_head.weight)
                nn.init.zeros_(self.binary_head.bias)
        if pre_process or post_process:
            self.setup_embeddings_and_output_layer()

    def set_input_tensor(self, t):
        self.encoder.set_input_tensor(t[0] if isinstance(t, (list, tuple)) else t)

    def forward(self, input_ids, attention_mask, tokentype_ids=None, lm_labels=None, inference_context=None):
        ext_mask = bert_extended_attention_mask(attention_mask)
        x = None
        if self.pre_process:
            x = self.embedding(input_ids=input_ids, position_ids=bert_position_ids(input_ids), tokentype_ids=tokentype_ids)
        hidden = self.encoder(hidden_states=x, attention_mask=ext_mask, inference_context=inference_context)
        if not self.post_process:
            return hidden
        binary_logits = None
        if self.add_binary_head:
            binary_logits = self.binary_head(self.pooler(hidden, 0))
        if self.return_embeddings:
            emb = hidden.transpose(0, 1)
            mask = attention_mask.unsqueeze(-1).to(emb.dtype)
            return (emb * mask).sum(1) / mask.sum(1).clamp_min(1)
        h = self.lm_head(hidden)
        w = self.shared_embedding_or_output_weight() if self.share_embeddings_and_output_weights else None
        logits, _ = self.output_layer(h, weight=w)
        if lm_labels is None:
            return logits.transpose(0, 1).contiguous(), binary_logits
        return self.compute_language_model_loss(lm_labels, logits), binary_logits
