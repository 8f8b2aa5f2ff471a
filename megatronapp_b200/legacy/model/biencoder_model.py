"""Dual-encoder retriever (reference ``megatron/legacy/model/biencoder_model.py``: ``BiEncoderModel`` /
``PretrainedBertModel``) for inverse-cloze-task (ICT) pre-training and open-retrieval QA (ORQA/REALM).

A query tower and a context tower (optionally one shared tower) each map ``[b, s]`` tokens to one vector: the
encoder's [CLS] hidden state, optionally projected to ``biencoder_projection_dim``.  Training uses in-batch
negatives over the whole data-parallel batch (``ict_loss``).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...models.bert.bert_model import bert_extended_attention_mask, bert_position_ids
from ...models.common.embeddings import LanguageModelEmbedding
from ...transformer.module import MegatronModule
from ...transformer.transformer_block import TransformerBlock


class PretrainedBertModel(MegatronModule):
    """BERT trunk + [CLS] pooling (+ linear projection); no LM / NSP heads."""

    def __init__(self, config, vocab_size: int, max_sequence_length: int, num_tokentypes: int = 2,
                 projection_dim: int = 0, transformer_layer_spec=None, pre_process=True, post_process=True):
        super().__init__(config)
        from ...models.bert.bert_model import bert_layer_local_spec

        self.pre_process, self.post_process = pre_process, post_process
        if pre_process:
            self.embedding = LanguageModelEmbedding(config, vocab_size, max_sequence_length, "learned_absolute",
                                                    num_tokentypes=num_tokentypes)
        self.encoder = TransformerBlock(config, transformer_layer_spec or bert_layer_local_spec(), pre_process=pre_process,
                                        post_process=post_process)
        self.projection_enc = None
        if post_process and projection_dim > 0:
            self.projection_enc = nn.Linear(config.hidden_size, projection_dim, dtype=config.params_dtype)
            config.init_method(self.projection_enc.weight)
            nn.init.zeros_(self.projection_enc.bias)

    def set_input_tensor(self, t):
        self.encoder.set_input_tensor(t)

    def forward(self, input_ids, attention_mask, tokentype_ids=None):
        x = None
        if self.pre_process:
            x = self.embedding(input_ids=input_ids, position_ids=bert_position_ids(input_ids), tokentype_ids=tokentype_ids)
        h = self.encoder(hidden_states=x, attention_mask=bert_extended_attention_mask(attention_mask))     # [s, b, h]
        if not self.post_process:
            return h
        pooled = h[0]
        return self.projection_enc(pooled) if self.projection_enc is not None else pooled

    def init_from_bert(self, bert_state_dict: dict):
        """Warm start from a pre-trained ``BertModel`` (``--bert-load``): embedding + encoder only."""
        own = self.state_dict()
        take = {k: v for k, v in bert_state_dict.items() if k in own and own[k].shape == v.shape}
        self.load_state_dict(take, strict=False)
        return sorted(take)


class BiEncoderModel(MegatronModule):
    def __init__(self, config, vocab_size: int, max_sequence_length: int, num_tokentypes: int = 2,
                 biencoder_projection_dim: int = 0, biencoder_shared_query_context_model: bool = False,
                 only_query_model: bool = False, only_context_model: bool = False, pre_process=True, post_process=True):
        super().__init__(config)
        assert not (only_query_model and only_context_model)
        self.shared = biencoder_shared_query_context_model
        mk = lambda: PretrainedBertModel(config, vocab_size, max_sequence_length, num_tokentypes, biencoder_projection_dim,
                                         pre_process=pre_process, post_process=post_process)
        self.use_query_model, self.use_context_model = not only_context_model, not only_query_model
        if self.shared:
            self.model = mk()
            self.query_model = self.context_model = None
        else:
            self.query_model = mk() if self.use_query_model else None
            self.context_model = mk() if self.use_context_model else None

    def set_input_tensor(self, t):
        for m in {self._tower(True), self._tower(False)} - {None}:
            m.set_input_tensor(t)

    def _tower(self, query: bool):
        return self.model if self.shared else (self.query_model if query else self.context_model)

    def embed_text(self, tokens, mask, types=None, query: bool = True):
        return self._tower(query)(tokens, mask, types)            # mask: [b, s] padding mask (1 = token)

    def forward(self, query_tokens, query_attention_mask, query_types, context_tokens, context_attention_mask,
                context_types):
        q = self.embed_text(query_tokens, query_attention_mask, query_types, True) if self.use_query_model else None
        c = self.embed_text(context_tokens, context_attention_mask, context_types, False) if self.use_context_model else None
        return q, c

    def init_state_dict_from_bert(self, bert_load=None):
        """Start both towers from a pre-trained BERT checkpoint (``--bert-load``) on iteration 0 of ICT / retriever training
        (reference biencoder_model.py:180-245).  Embedding and encoder are taken over, the LM / NSP heads are dropped; the
        context tower also starts from the query tower's projection so both embed into the same space."""
        from ...models.bert.bert_model import BertModel, bert_layer_local_spec
        from ...training.checkpointing import load_checkpoint
        from ...training.global_vars import get_args

        args = get_args()
        bert_load = bert_load or getattr(args, "bert_load", None)
        if bert_load is None:
            return []
        # a BERT of the towers' architecture on this rank's tensor-parallel slice: whatever format / layout the checkpoint has,
        # the regular loader reshards it; only embedding + encoder are kept
        any_tower = self._tower(True) or self._tower(False)
        emb = any_tower.embedding
        bert = BertModel(config=self.config, num_tokentypes=2 if getattr(args, "bert_binary_head", True) else 0,
                         transformer_layer_spec=bert_layer_local_spec(), vocab_size=args.padded_vocab_size,
                         max_sequence_length=emb.position_embeddings.weight.shape[0], add_binary_head=getattr(args, "bert_binary_head", True),
                         parallel_output=True)
        keep = (args.finetune, args.no_load_rng, args.bert_load)
        args.finetune, args.no_load_rng, args.bert_load = True, True, bert_load
        try:
            load_checkpoint([bert], None, None, load_arg="bert_load", strict=False)
        finally:
            args.finetune, args.no_load_rng, args.bert_load = keep
        dev = next(any_tower.parameters())
        bert_sd = {k: v.to(device=dev.device, dtype=dev.dtype) if torch.is_tensor(v) and v.is_floating_point() else v
                   for k, v in bert.state_dict().items()}
        taken = []
        for tower in {self._tower(True), self._tower(False)} - {None}:
            taken = tower.init_from_bert(bert_sd)
            if not taken:
                raise RuntimeError(f"no tensor of the BERT checkpoint in {bert_load} matches the bi-encoder towers")
        if not self.shared and self.query_model is not None and self.context_model is not None \
                and self.query_model.projection_enc is not None:
            self.context_model.projection_enc.load_state_dict(self.query_model.projection_enc.state_dict())
        return taken

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        names = ("model",) if self.shared else ("query_model", "context_model")
        return {n: getattr(self, n).state_dict(keep_vars=keep_vars) for n in names if getattr(self, n) is not None}

    def load_state_dict(self, state_dict, strict=True):
        if any(k in state_dict for k in ("model", "query_model", "context_model")) and \
                all(isinstance(v, dict) for k, v in state_dict.items() if k in ("model", "query_model", "context_model")):
            for n, sd in state_dict.items():
                if getattr(self, n, None) is not None:
                    getattr(self, n).load_state_dict(sd, strict=strict)
            return
        return super().load_state_dict(state_dict, strict=strict)


class _AllGatherBatch(torch.autograd.Function):
    """Concatenate the per-rank batches over the data-parallel group; backward keeps this rank's slice."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        world = torch.distributed.get_world_size(group)
        out = [torch.empty_like(x) for _ in range(world)]
        torch.distributed.all_gather(out, x.contiguous(), group=group)
        out[torch.distributed.get_rank(group)] = x
        return torch.cat(out, 0)

    @staticmethod
    def backward(ctx, g):
        world, rank = torch.distributed.get_world_size(ctx.group), torch.distributed.get_rank(ctx.group)
        return g.chunk(world, 0)[rank].contiguous(), None


def ict_loss(query_logits: torch.Tensor, context_logits: torch.Tensor, hidden_size: int, score_scaling: bool = False,
             topk: Sequence[int] = (1, 5), dp_group=None):
    """In-batch-negative retrieval loss (reference ``pretrain_ict.py loss_func``): every query must rank its own
    context first among the contexts of the WHOLE data-parallel batch.  Returns ``(loss, {top-k accuracies in %})``."""
    if dp_group is not None and torch.distributed.is_initialized() and torch.distributed.get_world_size(dp_group) > 1:
        query_logits, context_logits = _AllGatherBatch.apply(query_logits, dp_group), _AllGatherBatch.apply(context_logits, dp_group)
    scores = query_logits.float() @ context_logits.float().t()
    if score_scaling:
        scores = scores / math.sqrt(hidden_size)
    logp = F.log_softmax(scores, dim=1)
    n = scores.shape[0]
    labels = torch.arange(n, device=scores.device)
    loss = F.nll_loss(logp, labels)
    rank = (logp > logp.gather(1, labels[:, None])).sum(1)                 # how many contexts beat the true one
    stats = {f"top{k}_acc": (rank < min(k, n)).float().mean() * 100 for k in topk}
    return loss, stats
