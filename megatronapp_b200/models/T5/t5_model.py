"""T5 encoder-decoder (reference ``core/models/T5/t5_model.py``: shared embedding, encoder stack with
padding mask, decoder stack with causal self-attention + cross-attention to the encoder output,
tied LM head ``T5LMHead``; encoder/decoder pipeline split is expressed by ``add_encoder`` /
``add_decoder``)."""
from __future__ import annotations

import torch

from ...parallel.tp_layers import ColumnParallelLinear, RowParallelLinear
from ...transformer.attention import (CrossAttention, CrossAttentionSubmodules, DotProductAttention, SelfAttention,
                                      SelfAttentionSubmodules)
from ...transformer.mlp import MLP, MLPSubmodules
from ...transformer.norms import FusedLayerNorm
from ...transformer.spec_utils import ModuleSpec
from ...transformer.transformer_block import TransformerBlock, TransformerBlockSubmodules
from ...transformer.transformer_layer import TransformerLayer, TransformerLayerSubmodules
from ..common.embeddings import LanguageModelEmbedding
from ..common.language_module import LanguageModule


def _mlp():
    return ModuleSpec(module=MLP, submodules=MLPSubmodules(linear_fc1=ColumnParallelLinear, linear_fc2=RowParallelLinear))


def encoder_layer_spec() -> ModuleSpec:
    return ModuleSpec(module=TransformerLayer, submodules=TransformerLayerSubmodules(
        input_layernorm=FusedLayerNorm,
        self_attention=ModuleSpec(module=SelfAttention, params={"attn_mask_type": "padding"},
                                  submodules=SelfAttentionSubmodules(linear_qkv=ColumnParallelLinear,
                                                                     core_attention=DotProductAttention,
                                                                     linear_proj=RowParallelLinear)),
        pre_mlp_layernorm=FusedLayerNorm, mlp=_mlp()))


def decoder_layer_spec() -> ModuleSpec:
    return ModuleSpec(module=TransformerLayer, submodules=TransformerLayerSubmodules(
        input_layernorm=FusedLayerNorm,
        self_attention=ModuleSpec(module=SelfAttention, params={"attn_mask_type": "causal"},
                                  submodules=SelfAttentionSubmodules(linear_qkv=ColumnParallelLinear,
                                                                     core_attention=DotProductAttention,
                                                                     linear_proj=RowParallelLinear)),
        pre_cross_attn_layernorm=FusedLayerNorm,
        cross_attention=ModuleSpec(module=CrossAttention, params={"attn_mask_type": "padding"},
                                   submodules=CrossAttentionSubmodules(linear_q=ColumnParallelLinear,
                                                                       linear_kv=ColumnParallelLinear,
                                                                       core_attention=DotProductAttention,
                                                                       linear_proj=RowParallelLinear)),
        pre_mlp_layernorm=FusedLayerNorm, mlp=_mlp()))


def get_t5_encoder_block_spec(n):
    return TransformerBlockSubmodules(layer_specs=[encoder_layer_spec()] * n, layer_norm=FusedLayerNorm)


def get_t5_decoder_block_spec(n):
    return TransformerBlockSubmodules(layer_specs=[decoder_layer_spec()] * n, layer_norm=FusedLayerNorm)


def t5_extended_attention_mask(masks):
    return [None if m is None else (m < 0.5).unsqueeze(1) for m in masks]


class T5Model(LanguageModule):
    def __init__(self, config, encoder_config, transformer_encoder_layer_spec, transformer_decoder_layer_spec,
                 vocab_size: int, max_sequence_length: int, pre_process=True, post_process=True,
                 fp16_lm_cross_entropy=False, parallel_output=True, share_embeddings_and_output_weights=True,
                 position_embedding_type="learned_absolute", add_encoder=True, add_decoder=True,
                 relative_attention_num_buckets: int = 32, relative_attention_max_distance: int = 128):
        super().__init__(config)
        self.log_constructor_args(locals())
        # under a pipeline split, ``post_process`` of an encoder-only stage means "apply the encoder's final norm";
        # the LM head (and the tied-embedding bookkeeping of LanguageModule) exists on decoder stages only
        enc_post = post_process
        post_process = post_process and add_decoder
        self.pre_process, self.post_process = pre_process, post_process
        self.add_encoder, self.add_decoder = add_encoder, add_decoder
        self.share_embeddings_and_output_weights = share_embeddings_and_output_weights
        self.encoder_hidden_state = None
        if pre_process:
            self.embedding = LanguageModelEmbedding(config, vocab_size, max_sequence_length, position_embedding_type)
        self.position_embedding_type = position_embedding_type
        if position_embedding_type == "relative":       # reference t5_model.py: one bias table per stack
            from ..common.embeddings import RelativePositionEmbedding

            if add_encoder:
                self.encoder_relative_pos_emb = RelativePositionEmbedding(
                    True, config.init_method, config.num_attention_heads, relative_attention_num_buckets,
                    relative_attention_max_distance)
            if add_decoder:
                self.decoder_relative_pos_emb = RelativePositionEmbedding(
                    False, config.init_method, config.num_attention_heads, relative_attention_num_buckets,
                    relative_attention_max_distance)
        if add_encoder:
            self.encoder = TransformerBlock(encoder_config or config, transformer_encoder_layer_spec,
                                            pre_process=pre_process, post_process=True if add_decoder else enc_post)
        if add_decoder:
            self.decoder = TransformerBlock(config, transformer_decoder_layer_spec, pre_process=pre_process,
                                            post_process=post_process)
        self._renumber_layers()
        if post_process:
            self.output_layer = ColumnParallelLinear(config.hidden_size, vocab_size, config=config,
                                                     init_method=config.init_method, bias=share_embeddings_and_output_weights,
                                                     skip_bias_add=False, gather_output=not parallel_output,
                                                     skip_weight_param_allocation=pre_process and share_embeddings_and_output_weights)
        if pre_process or post_process:
            self.setup_embeddings_and_output_layer()

    def _renumber_layers(self):
        """Layer numbers (MegaScope hooks, checkpoint keys) count within the encoder resp. decoder stack.  The generic
        pipeline offset assumes one stack split evenly over all stages, which is wrong under the encoder/decoder split."""
        from ...parallel import state as ps

        pp, split = ps.get_pipeline_model_parallel_world_size(), ps.get_pipeline_model_parallel_split_rank()
        r = ps.get_pipeline_model_parallel_rank()
        for name, stage in (("encoder", r), ("decoder", r - (split or 0))):
            block = getattr(self, name, None)
            if block is None:
                continue
            base = stage * len(block.layers) if (pp > 1 and split is not None) else 0
            for i, layer in enumerate(block.layers):
                old, new = layer.layer_number, base + i + 1
                for mod in layer.modules():
                    if getattr(mod, "layer_number", None) == old:
                        mod.layer_number = new

    def set_input_tensor(self, t):
        t = t if isinstance(t, (list, tuple)) else [t]
        if self.add_encoder and self.add_decoder:
            self.encoder.set_input_tensor(t[0])
        elif self.add_encoder:
            self.encoder.set_input_tensor(t[0])
        elif self.add_decoder:
            if len(t) == 2:
                self.decoder.set_input_tensor(t[0])
                self.encoder_hidden_state = t[1]
            else:
                self.decoder.set_input_tensor(None)
                self.encoder_hidden_state = t[0]

    def forward(self, encoder_input_ids, decoder_input_ids, encoder_attn_mask, decoder_attn_mask,
                encoder_decoder_attn_mask, lm_labels=None, encoder_hidden_states=None, output_encoder_hidden_only=False,
                inference_context=None):
        def pos(ids):
            return torch.arange(ids.size(1), device=ids.device).unsqueeze(0).expand_as(ids)

        enc_mask, dec_mask, ed_mask = t5_extended_attention_mask([encoder_attn_mask, decoder_attn_mask,
                                                                  encoder_decoder_attn_mask])
        if self.add_encoder and encoder_hidden_states is None:
            x = self.embedding(encoder_input_ids, pos(encoder_input_ids)) if self.pre_process else None
            enc_bias = None
            if self.position_embedding_type == "relative":
                n = self.encoder_relative_pos_emb.get_relative_seq_len(None, self.encoder, x, self.config)
                enc_bias = self.encoder_relative_pos_emb(n, n)
            enc = self.encoder(hidden_states=x, attention_mask=enc_mask, attention_bias=enc_bias)
        else:
            enc = encoder_hidden_states if encoder_hidden_states is not None else self.encoder_hidden_state
        if not self.add_decoder or output_encoder_hidden_only:
            return enc
        y = self.embedding(decoder_input_ids, pos(decoder_input_ids)) if self.pre_process else None
        dec_bias = None
        if self.position_embedding_type == "relative":
            n = self.decoder_relative_pos_emb.get_relative_seq_len(inference_context, self.decoder, y, self.config)
            dec_bias = self.decoder_relative_pos_emb(n, n)
        dec = self.decoder(hidden_states=y, attention_mask=dec_mask, context=enc, context_mask=ed_mask,
                           attention_bias=dec_bias, inference_context=inference_context)
        if not self.post_process:
            return dec
        w = self.shared_embedding_or_output_weight() if self.share_embeddings_and_output_weights else None
        logits, _ = self.output_layer(dec, weight=w)
        if lm_labels is None:
            return logits.transpose(0, 1).contiguous()
        return self.compute_language_model_loss(lm_labels, logits)


# -- spec / helper names of the reference's ``T5/t5_spec.py`` and ``t5_model.py`` ----------------------------------------------------
def encoder_model_with_local_spec() -> ModuleSpec:
    return encoder_layer_spec()


def decoder_model_with_local_spec() -> ModuleSpec:
    return decoder_layer_spec()


encoder_model_with_transformer_engine_default_spec = encoder_model_with_local_spec     # one (native) implementation of the layers
decoder_model_with_transformer_engine_default_spec = decoder_model_with_local_spec


def get_t5_encoder_with_local_block_spec(num_layers: int):
    return get_t5_encoder_block_spec(num_layers)


def get_t5_decoder_with_local_block_spec(num_layers: int):
    return get_t5_decoder_block_spec(num_layers)


get_t5_encoder_with_transformer_engine_block_spec = get_t5_encoder_with_local_block_spec
get_t5_decoder_with_transformer_engine_block_spec = get_t5_decoder_with_local_block_spec


def t5_position_ids(token_ids: torch.Tensor) -> torch.Tensor:
    """``[b, s]`` position ids ``0 .. s-1`` for every row."""
    return torch.arange(token_ids.size(1), dtype=torch.long, device=token_ids.device).unsqueeze(0).expand_as(token_ids)
