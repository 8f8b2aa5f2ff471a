"""Legacy model classes (reference ``megatron/legacy/model/{gpt_model,bert_model,t5_model}.py``).

The reference keeps a second, older implementation of every model (``ParallelTransformer``).  Here the legacy
classes are thin shells over the core models: same forward signatures as the legacy classes, and
``state_dict_for_save_checkpoint`` / ``load_state_dict`` speak the legacy nested naming so that
``--use-legacy-models`` runs and legacy checkpoints load -- without a second transformer implementation to keep fast.
"""
from __future__ import annotations

import torch

from ...transformer.module import MegatronModule
from ..state_dict import core_to_legacy, legacy_to_core


class _LegacyShell(MegatronModule):
    _stack = "decoder"

    def __init__(self, core):
        super().__init__(core.config)
        self.language_model = core              # attribute name the legacy tools reach for

    # pipeline / DDP plumbing delegates to the core model
    def set_input_tensor(self, t):
        self.language_model.set_input_tensor(t)

    def shared_embedding_or_output_weight(self):
        return self.language_model.shared_embedding_or_output_weight()

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("language_model"), name)

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        return core_to_legacy(self.language_model.state_dict(keep_vars=keep_vars), self._stack)

    def load_state_dict(self, state_dict, strict=True):
        if "language_model" in state_dict and isinstance(state_dict["language_model"], dict):
            state_dict = legacy_to_core(state_dict, self._stack)
        elif any(k.startswith("language_model.") for k in state_dict):
            state_dict = {k[len("language_model."):]: v for k, v in state_dict.items()}
        return self.language_model.load_state_dict(state_dict, strict=strict)

    def sharded_state_dict(self, prefix="", *a, **k):
        return self.language_model.sharded_state_dict(prefix, *a, **k)


class GPTModel(_LegacyShell):
    """``GPTModel(config, num_tokentypes=0, parallel_output=True, pre_process=True, post_process=True)``."""

    def __init__(self, config, num_tokentypes=0, parallel_output=True, pre_process=True, post_process=True, args=None):
        from ...models.gpt.gpt_layer_specs import get_gpt_layer_local_spec
        from ...models.gpt.gpt_model import GPTModel as CoreGPT
        from ...training.global_vars import get_args

        args = args or get_args()
        rope = getattr(args, "position_embedding_type", "learned_absolute")
        core = CoreGPT(config, get_gpt_layer_local_spec(), args.padded_vocab_size, args.max_position_embeddings,
                       pre_process=pre_process, post_process=post_process, parallel_output=parallel_output,
                       share_embeddings_and_output_weights=not getattr(args, "untie_embeddings_and_output_weights", False),
                       position_embedding_type=rope, rotary_percent=getattr(args, "rotary_percent", 1.0),
                       rotary_base=getattr(args, "rotary_base", 10000))
        super().__init__(core)

    def forward(self, input_ids, position_ids, attention_mask, retriever_input_ids=None, retriever_position_ids=None,
                retriever_attn_mask=None, labels=None, tokentype_ids=None, inference_context=None,
                inference_params=None):
        return self.language_model(input_ids, position_ids, attention_mask, labels=labels,
                                   inference_context=inference_context or inference_params)


class BertModel(_LegacyShell):
    _stack = "encoder"

    def __init__(self, config, num_tokentypes=2, add_binary_head=True, parallel_output=True, pre_process=True,
                 post_process=True, args=None):
        from ...models.bert.bert_model import BertModel as CoreBert, bert_layer_local_spec
        from ...training.global_vars import get_args

        args = args or get_args()
        core = CoreBert(config, num_tokentypes, bert_layer_local_spec(), args.padded_vocab_size,
                        args.max_position_embeddings, pre_process=pre_process, post_process=post_process,
                        parallel_output=parallel_output, add_binary_head=add_binary_head)
        super().__init__(core)

    def forward(self, bert_model_input, attention_mask, tokentype_ids=None, lm_labels=None):
        return self.language_model(bert_model_input, attention_mask, tokentype_ids=tokentype_ids, lm_labels=lm_labels)


class T5Model(_LegacyShell):
    def __init__(self, config, num_tokentypes=0, parallel_output=True, pre_process=True, post_process=True,
                 add_encoder=True, add_decoder=True, args=None):
        from ...models.T5 import T5Model as CoreT5, get_t5_decoder_block_spec, get_t5_encoder_block_spec
        from ...training.global_vars import get_args

        args = args or get_args()
        n_enc = getattr(args, "encoder_num_layers", None) or args.num_layers
        n_dec = getattr(args, "decoder_num_layers", None) or args.num_layers
        core = CoreT5(config, config, get_t5_encoder_block_spec(n_enc), get_t5_decoder_block_spec(n_dec),
                      vocab_size=args.padded_vocab_size, max_sequence_length=args.max_position_embeddings,
                      pre_process=pre_process, post_process=post_process, parallel_output=parallel_output,
                      add_encoder=add_encoder, add_decoder=add_decoder)
        super().__init__(core)

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):     # two stacks: keep the flat core names
        return {"language_model": dict(self.language_model.state_dict(keep_vars=keep_vars))}

    def load_state_dict(self, state_dict, strict=True):
        return self.language_model.load_state_dict(state_dict.get("language_model", state_dict), strict=strict)

    def forward(self, encoder_input_ids, decoder_input_ids, encoder_attn_mask, decoder_attn_mask,
                encoder_decoder_attn_mask, tokentype_ids=None, lm_labels=None, enc_hidden_states=None):
        return self.language_model(encoder_input_ids, decoder_input_ids, encoder_attn_mask, decoder_attn_mask,
                                   encoder_decoder_attn_mask, lm_labels=lm_labels, encoder_hidden_states=enc_hidden_states)
