"""Evidence indexer (reference ``legacy/indexer.py`` ``IndexBuilder``): embed every evidence passage with the bi-encoder's
context tower -- the data-parallel ranks split the passages -- and write the merged embedding store that the retrieval
evaluation (``tasks/orqa``) and the REALM-style retriever load.

    builder = IndexBuilder(model, tokenizer, evidence_data_path, embedding_path, seq_length=256, batch_size=128)
    builder.build_and_save_index()
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .data.orqa_wiki_dataset import get_open_retrieval_batch, get_open_retrieval_wiki_dataset
from .data.realm_index import OpenRetreivalDataStore


class IndexBuilder:
    def __init__(self, model, tokenizer, evidence_data_path: str, embedding_path: str, seq_length: int = 256, batch_size: int = 128,
                 log_interval: int = 1000, sample_rate: float = 1.0):
        self.model = model[0] if isinstance(model, (list, tuple)) else model
        self.batch_size, self.log_interval = batch_size, log_interval
        self.rank, self.world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
        self.is_main_builder = self.rank == 0
        self.dataset = get_open_retrieval_wiki_dataset(evidence_data_path, tokenizer, seq_length, sample_rate)
        self.evidence_embedder_obj = OpenRetreivalDataStore(embedding_path, load_from_path=False, rank=self.rank)
        self.iteration = self.total_processed = 0

    def track_and_report_progress(self, batch_size: int):
        self.iteration += 1
        self.total_processed += batch_size * self.world
        if self.is_main_builder and self.iteration % self.log_interval == 0:
            print(f"Batch {self.iteration:10d} | Total {self.total_processed:10d}", flush=True)

    @torch.no_grad()
    def build_and_save_index(self):
        """This rank embeds rows ``rank::world``; shards are saved, rank 0 merges them into ``embedding_path``."""
        model = self.model
        was_training = model.training
        model.eval()
        dev = next(model.parameters()).device
        mine = list(range(self.rank, len(self.dataset), self.world))
        for b0 in range(0, len(mine), self.batch_size):
            rows = [self.dataset[i] for i in mine[b0:b0 + self.batch_size]]
            row_id, ctx, mask, types, _ = get_open_retrieval_batch(rows)
            emb = model.embed_text(ctx.to(dev), mask.to(dev), types.to(dev), query=False)
            self.evidence_embedder_obj.add_block_data(row_id.numpy(), emb.float().cpu().numpy())
            self.track_and_report_progress(len(rows))
        self.evidence_embedder_obj.save_shard()
        if dist.is_initialized():
            dist.barrier()
        if self.is_main_builder:
            self.evidence_embedder_obj.merge_shards_and_save()
            self.evidence_embedder_obj.clear()                  # the merged store lives on disk
        if dist.is_initialized():
            dist.barrier()
        model.train(was_training)
