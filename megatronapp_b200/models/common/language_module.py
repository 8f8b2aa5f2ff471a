"""Shared behaviour of language models: loss computation and tied-embedding setup.

Parity: reference ``models/common/language_module/language_module.py``
(``compute_language_model_loss``, ``setup_embeddings_and_output_layer``,
``shared_embedding_or_output_weight``).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ...parallel import state as ps
from ...parallel.cross_entropy import vocab_parallel_cross_entropy
from ...transformer.module import MegatronModule


class LanguageModule(MegatronModule):
    def compute_language_model_loss(self, labels: torch.Tensor, logits: torch.Tensor) -> torch.Tensor:
        # labels [b, s] -> [s, b]; logits [s, b, V/tp]
        labels = labels.transpose(0, 1).contiguous()
        loss = vocab_parallel_cross_entropy(logits, labels)
        return loss.transpose(0, 1).contiguous()  # [b, s]

    def setup_embeddings_and_output_layer(self) -> None:
        if self.pre_process:
            self.embedding.word_embeddings.weight.is_embedding_or_output_parameter = True
        if self.post_process and getattr(self, "output_layer", None) is not None and self.output_layer.weight is not None:
            self.output_layer.weight.is_embedding_or_output_parameter = True
        if not self.share_embeddings_and_output_weights:
            return
        if ps.get_pipeline_model_parallel_world_size() == 1:
            self.shared_embedding_or_output_weight().zero_out_wgrad = True
            return
        if ps.is_pipeline_first_stage() and self.pre_process and not self.post_process:
            self.shared_embedding_or_output_weight().shared_embedding = True
        if self.pre_process and not ps.is_pipeline_first_stage():
            # first decoder stage of an encoder-decoder split: a second copy of the input embedding, zeroed so that
            # the all-reduce below leaves every copy equal to the first stage's initialisation
            self.embedding.word_embeddings.weight.data.fill_(0)
            self.embedding.word_embeddings.weight.shared = True
            self.embedding.word_embeddings.weight.shared_embedding = True
            if getattr(self.embedding, "position_embeddings", None) is not None:
                self.embedding.position_embeddings.weight.data.fill_(0)
        if self.post_process and not self.pre_process:
            # last stage keeps its own copy of the word embedding for the output layer,
            # initialised to zero and filled by an all-reduce with the first stage
            self.output_layer.weight.data.fill_(0)
            self.output_layer.weight.shared = True
            self.output_layer.weight.shared_embedding = True
        if dist.is_initialized() and ps.is_rank_in_embedding_group():
            w = self.shared_embedding_or_output_weight()
            g = ps.get_embedding_group()
            if g is not None and not isinstance(g, ps._SingleGroup):
                dist.all_reduce(w.data, group=g)
        if dist.is_initialized() and ps.get_pipeline_model_parallel_split_rank() is not None and self.pre_process \
                and ps.is_rank_in_position_embedding_group() and getattr(self.embedding, "position_embeddings", None) is not None:
            g = ps.get_position_embedding_group()
            if g is not None and not isinstance(g, ps._SingleGroup):
                dist.all_reduce(self.embedding.position_embeddings.weight.data, group=g)

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        """Tied embeddings: the last pipeline stage's private copy of the word embedding is stored under the
        EMBEDDING's key as a non-main replica, so one tensor lands in the checkpoint and any PP layout can load it
        (reference ``LanguageModule.tie_embeddings_and_output_weights_state_dict``)."""
        import dataclasses

        sd = super().sharded_state_dict(prefix, sharded_offsets, metadata)
        k_out, k_emb = f"{prefix}output_layer.weight", f"{prefix}embedding.word_embeddings.weight"
        if getattr(self, "share_embeddings_and_output_weights", False) and self.post_process and not self.pre_process \
                and k_out in sd:
            t = sd.pop(k_out)
            rid = t.replica_id if isinstance(t.replica_id, tuple) else (t.replica_id,)
            sd[k_out] = dataclasses.replace(t, key=k_emb, replica_id=(1,) + tuple(rid[1:]), allow_shape_mismatch=True)
        return sd

    def shared_embedding_or_output_weight(self):
        if self.pre_process:
            return self.embedding.word_embeddings.weight
        if self.post_process:
            return self.output_layer.weight
        return None
