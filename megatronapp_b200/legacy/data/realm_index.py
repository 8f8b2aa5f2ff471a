"""Evidence embedding store + maximum-inner-product index for open-retrieval QA (reference ``legacy/data/realm_index.py``:
``OpenRetreivalDataStore`` (sic) and ``FaissMIPSIndex``).

The store maps an evidence row id to its embedding; every data-parallel rank fills and saves its own shard
(``<path>_tmp/<rank>.pkl``), rank 0 merges them into ``<path>``.  The index answers top-k inner-product queries; faiss is
not part of this stack, so it is backed by ``data/retro/index.py`` (exact GEMM search by default, IVF-PQ via ``index_str``)."""
from __future__ import annotations

import os
import pickle
import shutil
from typing import Dict, Optional, Tuple

import numpy as np
import torch


class OpenRetreivalDataStore:
    def __init__(self, embedding_path: Optional[str] = None, load_from_path: bool = True, rank: Optional[int] = None):
        self.embed_data: Dict[int, np.ndarray] = {}
        self.embedding_path = embedding_path
        self.rank = rank if rank is not None else (torch.distributed.get_rank() if torch.distributed.is_initialized() else 0)
        if load_from_path and embedding_path and os.path.exists(embedding_path):
            self.load_from_file()
        self.temp_dir_name = (os.path.splitext(embedding_path)[0] + "_tmp") if embedding_path else None

    def state(self):
        return {"embed_data": self.embed_data}

    def clear(self):
        self.embed_data = {}

    def load_from_file(self):
        with open(self.embedding_path, "rb") as f:
            self.embed_data = pickle.load(f)["embed_data"]

    def add_block_data(self, row_id, block_embeds, allow_overwrite: bool = False):
        block_embeds = np.asarray(block_embeds, dtype=np.float32)
        for idx, emb in zip(np.asarray(row_id).reshape(-1).tolist(), block_embeds):
            if not allow_overwrite and idx in self.embed_data:
                raise ValueError(f"evidence row {idx} was embedded twice")
            self.embed_data[int(idx)] = np.float16(emb)

    def save_shard(self):
        os.makedirs(self.temp_dir_name, exist_ok=True)
        with open(os.path.join(self.temp_dir_name, f"{self.rank}.pkl"), "wb") as f:
            pickle.dump(self.state(), f)

    def merge_shards_and_save(self):
        """rank 0, after every rank's ``save_shard`` (callers barrier in between)"""
        seen = len(self.embed_data)
        for name in sorted(os.listdir(self.temp_dir_name)):
            if name == f"{self.rank}.pkl":
                continue
            with open(os.path.join(self.temp_dir_name, name), "rb") as f:
                data = pickle.load(f)["embed_data"]
            overlap = set(data) & set(self.embed_data)
            if overlap:
                raise ValueError(f"shards overlap on rows {sorted(overlap)[:5]}")
            self.embed_data.update(data)
        with open(self.embedding_path, "wb") as f:
            pickle.dump(self.state(), f)
        shutil.rmtree(self.temp_dir_name, ignore_errors=True)
        return len(self.embed_data) - seen


class FaissMIPSIndex:
    """Top-k inner-product search over the store's embeddings (name kept from the reference)."""

    def __init__(self, embed_size: int, embed_data: Optional[OpenRetreivalDataStore] = None, use_gpu: bool = False,
                 index_str: str = "Flat"):
        self.embed_size, self.embed_data, self.index_str = embed_size, embed_data, index_str
        self.device = "cuda" if (use_gpu and torch.cuda.is_available()) else "cpu"
        self.ids = np.zeros(0, dtype=np.int64)
        self._build()

    def _build(self):
        from ...data.retro.index import index_factory

        self.mips_index = index_factory(self.index_str, self.embed_size, self.device)
        self.ids = np.zeros(0, dtype=np.int64)

    def reset_index(self):
        self._build()
        if self.embed_data is not None:
            self.embed_data.clear()

    def update_index(self):
        self._build()
        if self.embed_data is not None:
            self.add_embed_data(self.embed_data)

    def add_embed_data(self, all_embed_data: OpenRetreivalDataStore):
        ids = np.fromiter(all_embed_data.embed_data.keys(), dtype=np.int64)
        if not len(ids):
            return
        x = np.stack([np.float32(all_embed_data.embed_data[int(i)]) for i in ids])
        if not self.mips_index.is_trained:
            self.mips_index.train(x)
        self.mips_index.add(x)
        self.ids = np.concatenate([self.ids, ids])
        all_embed_data.clear()                        # the index now owns the vectors (as in the reference)

    def search_mips_index(self, query_embeds, top_k: int, reconstruct: bool = True) -> Tuple[np.ndarray, np.ndarray]:
        """``reconstruct=False``: (scores, evidence ids); ``True``: (scores, embeddings of the hits) for exact indexes."""
        q = np.float32(torch.as_tensor(query_embeds).detach().float().cpu().numpy())
        scores, idx = self.mips_index.search(q, top_k)
        if reconstruct:
            if not hasattr(self.mips_index, "x"):
                raise NotImplementedError("reconstruction needs the exact (Flat) index")
            return scores, self.mips_index.x[torch.as_tensor(idx)].cpu().numpy()
        return scores, np.where(idx >= 0, self.ids[np.clip(idx, 0, max(len(self.ids) - 1, 0))], -1)
