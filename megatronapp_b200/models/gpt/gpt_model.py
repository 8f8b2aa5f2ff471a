"""GPT / Llama / Mixtral-style decoder-only language model.

Parity: reference ``models/gpt/gpt_model.py`` (constructor signature, ``forward`` :255-429
returning per-token loss when ``labels`` is given else ``[b, s, V/tp]`` logits,
``set_input_tensor``, tied embeddings, MegaScan "decoder" scope :333-349, parameter naming:
``embedding.word_embeddings.weight``, ``decoder.layers.N.*``, ``decoder.final_layernorm``,
``output_layer.weight``).
"""
from __future__ import annotations

from typing import Optional

import torch

from ...parallel import state as ps
from ...parallel.tp_layers import ColumnParallelLinear
from ...scan.tracer import scope as _scope
from ...transformer.spec_utils import ModuleSpec
from ...transformer.transformer_block import TransformerBlock
from ..common.embeddings import LanguageModelEmbedding, MultimodalRotaryEmbedding, RotaryEmbedding
from ..common.language_module import LanguageModule


class GPTModel(LanguageModule):
    def __init__(self, config, transformer_layer_spec, vocab_size: int, max_sequence_length: int,
                 pre_process: bool = True, post_process: bool = True, fp16_lm_cross_entropy: bool = False,
                 parallel_output: bool = True, share_embeddings_and_output_weights: bool = False,
                 position_embedding_type: str = "learned_absolute", rotary_percent: float = 1.0,
                 rotary_base: int = 10000, rope_scaling: bool = False, rope_scaling_factor: float = 8.0,
                 scatter_embedding_sequence_parallel: bool = True,
                 seq_len_interpolation_factor: Optional[float] = None, mtp_block_spec=None):
        super().__init__(config)
        self.log_constructor_args(locals())
        self.transformer_layer_spec = transformer_layer_spec
        self.vocab_size, self.max_sequence_length = vocab_size, max_sequence_length
        self.pre_process, self.post_process = pre_process, post_process
        self.fp16_lm_cross_entropy = fp16_lm_cross_entropy
        self.parallel_output = parallel_output
        self.share_embeddings_and_output_weights = share_embeddings_and_output_weights
        self.position_embedding_type = position_embedding_type
        self.model_type = "encoder_or_decoder"
        self.max_position_embeddings = max_sequence_length
        if pre_process:
            self.embedding = LanguageModelEmbedding(config, vocab_size, max_sequence_length,
                                                    position_embedding_type,
                                                    scatter_to_sequence_parallel=scatter_embedding_sequence_parallel)
        if position_embedding_type == "rope" and not config.multi_latent_attention:
            self.rotary_pos_emb = RotaryEmbedding(config.kv_channels, rotary_percent, config.rotary_interleaved,
                                                  seq_len_interpolation_factor, rotary_base, rope_scaling,
                                                  rope_scaling_factor)
        if position_embedding_type == "mrope":
            self.rotary_pos_emb = MultimodalRotaryEmbedding(config.kv_channels, rotary_percent, config.rotary_interleaved,
                                                            seq_len_interpolation_factor, rotary_base)
            self.mrope_section = list(getattr(config, "mrope_section", None) or [])
            if not self.mrope_section:
                raise ValueError("position_embedding_type 'mrope' needs config.mrope_section (--mrope-section t h w)")
        self.decoder = TransformerBlock(config, transformer_layer_spec, pre_process=pre_process,
                                        post_process=post_process)
        if post_process:
            self.output_layer = ColumnParallelLinear(
                config.hidden_size, vocab_size, config=config, init_method=config.init_method, bias=False,
                skip_bias_add=False, gather_output=not parallel_output,
                skip_weight_param_allocation=pre_process and share_embeddings_and_output_weights)
        self.mtp = None
        if post_process and (mtp_block_spec is not None or config.mtp_num_layers):
            from ...transformer.multi_token_prediction import MultiTokenPredictionBlock, get_mtp_block_spec

            spec = mtp_block_spec or get_mtp_block_spec(config, transformer_layer_spec if isinstance(
                transformer_layer_spec, ModuleSpec) else transformer_layer_spec.layer_specs[-1])
            self.mtp = MultiTokenPredictionBlock(config, spec)
            if not pre_process:   # the last stage needs its own embedding to feed the MTP depths
                self.embedding = LanguageModelEmbedding(config, vocab_size, max_sequence_length, position_embedding_type,
                                                        scatter_to_sequence_parallel=scatter_embedding_sequence_parallel)
        if pre_process or post_process:
            self.setup_embeddings_and_output_layer()

    def set_input_tensor(self, input_tensor):
        if isinstance(input_tensor, (list, tuple)):
            input_tensor = input_tensor[0]
        self.decoder.set_input_tensor(input_tensor)

    def forward(self, input_ids, position_ids, attention_mask=None, decoder_input=None, labels=None,
                inference_context=None, packed_seq_params=None, extra_block_kwargs=None,
                runtime_gather_output=None, inference_params=None, loss_mask=None):
        inference_context = inference_context or inference_params
        if decoder_input is None and self.pre_process:
            decoder_input = self.embedding(input_ids=input_ids, position_ids=position_ids)
        rotary = None
        if self.position_embedding_type == "rope" and not self.config.multi_latent_attention:
            s = self.rotary_pos_emb.get_rotary_seq_len(inference_context, self.decoder, decoder_input, self.config)
            rotary = self.rotary_pos_emb(s)
            if self.config.context_parallel_size > 1:
                from ...parallel.context_parallel import get_pos_emb_on_this_cp_rank

                rotary = get_pos_emb_on_this_cp_rank(rotary, 0)
        if self.position_embedding_type == "mrope":
            pid = position_ids if position_ids.dim() == 3 else position_ids.unsqueeze(0).expand(3, -1, -1)
            rotary = self.rotary_pos_emb(pid, self.mrope_section)                     # [s, b, rot] per-sample table
        with _scope("decoder"):
            hidden = self.decoder(hidden_states=decoder_input, attention_mask=attention_mask,
                                  inference_context=inference_context, rotary_pos_emb=rotary,
                                  packed_seq_params=packed_seq_params, **(extra_block_kwargs or {}))
        if not self.post_process:
            return hidden
        w = self.shared_embedding_or_output_weight() if self.share_embeddings_and_output_weights else None
        if self.mtp is not None and labels is not None and self.training:
            hidden = self.mtp(input_ids=input_ids, position_ids=position_ids, hidden_states=hidden,
                              attention_mask=attention_mask, labels=labels, loss_mask=loss_mask, rotary_pos_emb=rotary,
                              embedding=self.embedding, output_layer=self.output_layer, output_weight=w,
                              compute_loss=self.compute_language_model_loss)
        logits, _ = self.output_layer(hidden, weight=w, runtime_gather_output=runtime_gather_output)
        if labels is None:
            return logits.transpose(0, 1).contiguous()  # [s, b, V] -> [b, s, V]
        return self.compute_language_model_loss(labels, logits)

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        sd = super().sharded_state_dict(prefix, sharded_offsets, metadata)
        return sd
