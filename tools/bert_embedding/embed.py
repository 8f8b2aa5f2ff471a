"""Sentence/chunk embeddings from a BERT encoder (reference ``tools/bert_embedding/{embed,dataset}.py``): mean pooling
of the final hidden states over non-padding tokens, batched; used by the Retro database builder as ``embedder``."""
from __future__ import annotations

import os
import sys
from typing import List, Optional

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


class BertEmbedder:
    def __init__(self, model, tokenizer=None, max_seq_length: int = 256, batch_size: int = 128):
        self.model, self.tok, self.max_seq_length, self.batch_size = model, tokenizer, max_seq_length, batch_size

    @torch.no_grad()
    def embed_tokens(self, token_ids: torch.Tensor, pad_id: int = 0) -> torch.Tensor:
        """``token_ids [n, s]`` -> ``[n, hidden]`` (mean over non-pad positions of the encoder output)."""
        self.model.eval()
        dev = next(self.model.parameters()).device
        outs = []
        for i in range(0, token_ids.shape[0], self.batch_size):
            ids = token_ids[i:i + self.batch_size].to(dev)
            mask = (ids != pad_id).long()
            from megatronapp_b200.models.bert.bert_model import bert_extended_attention_mask, bert_position_ids

            x = self.model.embedding(input_ids=ids, position_ids=bert_position_ids(ids), tokentype_ids=torch.zeros_like(ids))
            h = self.model.encoder(hidden_states=x, attention_mask=bert_extended_attention_mask(mask))     # [s, b, h]
            h = h.transpose(0, 1).float()
            m = mask.unsqueeze(-1).float()
            outs.append((h * m).sum(1) / m.sum(1).clamp_min(1.0))
        return torch.cat(outs).cpu()

    def embed_text(self, texts: List[str]) -> torch.Tensor:
        assert self.tok is not None, "a tokenizer is required to embed raw text"
        rows = []
        for t in texts:
            ids = [self.tok.cls] + self.tok.tokenize(t)[: self.max_seq_length - 2] + [self.tok.sep]
            rows.append(ids + [self.tok.pad] * (self.max_seq_length - len(ids)))
        return self.embed_tokens(torch.tensor(rows), self.tok.pad)

    __call__ = embed_tokens
