"""Sequence classification / multiple choice heads on the BERT encoder.

Parity: reference ``megatron/legacy/model/classification.py`` (``Classification``: encoder + pooler ->
dropout -> dense ``num_classes``) and ``multiple_choice.py`` (``MultipleChoice``: ``[b, choices, s]`` inputs
flattened, one score per choice), used by ``tasks/glue`` and ``tasks/race``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ...transformer.transformer_block import TransformerBlock
from ..common.embeddings import LanguageModelEmbedding
from ..common.language_module import LanguageModule
from .bert_model import Pooler, bert_extended_attention_mask, bert_position_ids


class Classification(LanguageModule):
    def __init__(self, config, transformer_layer_spec, vocab_size: int, max_sequence_length: int, num_classes: int,
                 num_tokentypes: int = 2, pre_process=True, post_process=True):
        super().__init__(config)
        self.pre_process, self.post_process, self.num_classes = pre_process, post_process, num_classes
        self.share_embeddings_and_output_weights = False
        if pre_process:
            self.embedding = LanguageModelEmbedding(config, vocab_size, max_sequence_length, "learned_absolute",
                                                    num_tokentypes=num_tokentypes)
        self.encoder = TransformerBlock(config, transformer_layer_spec, pre_process=pre_process, post_process=post_process)
        if post_process:
            self.pooler = Pooler(config.hidden_size, config.init_method, config, config.sequence_parallel)
            self.classification_dropout = nn.Dropout(config.hidden_dropout)
            self.classification_head = nn.Linear(config.hidden_size, num_classes, dtype=config.params_dtype)
            config.init_method(self.classification_head.weight)
            nn.init.zeros_(self.classification_head.bias)

    def set_input_tensor(self, t):
        self.encoder.set_input_tensor(t[0] if isinstance(t, (list, tuple)) else t)

    def _encode(self, input_ids, attention_mask, tokentype_ids):
        x = None
        if self.pre_process:
            x = self.embedding(input_ids=input_ids, position_ids=bert_position_ids(input_ids), tokentype_ids=tokentype_ids)
        return self.encoder(hidden_states=x, attention_mask=bert_extended_attention_mask(attention_mask))

    def forward(self, input_ids, attention_mask, tokentype_ids=None):
        hidden = self._encode(input_ids, attention_mask, tokentype_ids)
        if not self.post_process:
            return hidden
        pooled = self.pooler(hidden, 0)
        return self.classification_head(self.classification_dropout(pooled)).view(-1, self.num_classes)

    def load_pretrained_encoder(self, bert_state_dict, strict: bool = False):
        """Initialise embedding + encoder (+ pooler) from a pretrained ``BertModel`` state dict."""
        own = self.state_dict()
        picked = {k: v for k, v in bert_state_dict.items() if k in own and own[k].shape == v.shape}
        return self.load_state_dict(picked, strict=strict)


class MultipleChoice(Classification):
    def __init__(self, config, transformer_layer_spec, vocab_size: int, max_sequence_length: int, num_tokentypes: int = 2,
                 pre_process=True, post_process=True):
        super().__init__(config, transformer_layer_spec, vocab_size, max_sequence_length, 1, num_tokentypes, pre_process,
                         post_process)

    def forward(self, input_ids, attention_mask, tokentype_ids=None):
        # [b, choices, s] -> [b * choices, s]
        nc = input_ids.shape[1]
        flat = lambda t: None if t is None else t.reshape(-1, t.shape[-1])
        hidden = self._encode(flat(input_ids), flat(attention_mask), flat(tokentype_ids))
        if not self.post_process:
            return hidden
        pooled = self.pooler(hidden, 0)
        return self.classification_head(self.classification_dropout(pooled)).view(-1, nc)
