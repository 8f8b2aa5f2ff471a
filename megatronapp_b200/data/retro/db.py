"""Chunk database (reference ``core/datasets/retro/db/{build,dataset,utils}.py``): every document of every corpus in the
blend is cut into ``chunk_length``-token chunks; a chunk is the row ``(dataset, document, offset, length)``.  The database
is what gets indexed and what neighbours are fetched from (chunk + continuation inside the same document)."""
from __future__ import annotations

import json
import os
from typing import List, Tuple

import numpy as np

from ..gpt_dataset import parse_blend
from ..indexed_dataset import IndexedDataset
from .config import RetroPreprocessingConfig


def build_db(cfg: RetroPreprocessingConfig) -> "ChunkDB":
    prefixes, _ = parse_blend(cfg.retro_gpt_data_path)
    rows: List[np.ndarray] = []
    for di, prefix in enumerate(prefixes):
        ds = IndexedDataset(prefix)
        lens = np.asarray(ds.sequence_lengths, dtype=np.int64)
        n_chunks = np.maximum(1, -(-np.maximum(lens - 1, 1) // cfg.retro_gpt_chunk_length))
        doc = np.repeat(np.arange(len(lens), dtype=np.int64), n_chunks)
        first = np.concatenate([[0], np.cumsum(n_chunks)[:-1]])
        off = (np.arange(int(n_chunks.sum()), dtype=np.int64) - np.repeat(first, n_chunks)) * cfg.retro_gpt_chunk_length
        ln = np.minimum(cfg.retro_gpt_chunk_length, lens[doc] - off)
        keep = ln > 0
        rows.append(np.stack([np.full(doc.shape, di, dtype=np.int64), doc, off, ln], 1)[keep])
    chunks = np.concatenate(rows) if rows else np.zeros((0, 4), dtype=np.int64)
    os.makedirs(cfg.path("db"), exist_ok=True)
    np.save(cfg.path("db", "chunks.npy"), chunks)
    rng = np.random.RandomState(cfg.retro_gpt_seed)
    n_train = min(len(chunks), int(cfg.retro_index_ntrain))
    np.save(cfg.path("db", "sampled.npy"), np.sort(rng.choice(len(chunks), n_train, replace=False)))
    with open(cfg.path("db", "meta.json"), "w") as f:
        json.dump({"prefixes": prefixes, "n_chunks": int(len(chunks)), "chunk_length": cfg.retro_gpt_chunk_length}, f)
    return ChunkDB(cfg)


class ChunkDB:
    def __init__(self, cfg: RetroPreprocessingConfig):
        self.cfg = cfg
        with open(cfg.path("db", "meta.json")) as f:
            self.meta = json.load(f)
        self.chunks = np.load(cfg.path("db", "chunks.npy"), mmap_mode="r")
        self.sampled = np.load(cfg.path("db", "sampled.npy"))
        self.datasets = [IndexedDataset(p) for p in self.meta["prefixes"]]

    def __len__(self) -> int:
        return self.chunks.shape[0]

    def doc_key(self, i: int) -> Tuple[int, int]:
        return int(self.chunks[i, 0]), int(self.chunks[i, 1])

    def tokens(self, i: int, length: int = None, pad: int = 0) -> np.ndarray:
        """``length`` tokens starting at chunk ``i`` (None: the chunk itself, padded to chunk_length); a longer window runs
        on into the following chunks of the same document -- the neighbour's *continuation*."""
        di, doc, off, ln = (int(x) for x in self.chunks[i])
        want = length or self.cfg.retro_gpt_chunk_length
        ds = self.datasets[di]
        have = min(want, int(ds.sequence_lengths[doc]) - off)
        t = np.asarray(ds.get(doc, off, have), dtype=np.int64)
        return t if have == want else np.concatenate([t, np.full(want - have, pad, dtype=np.int64)])

    def batch_tokens(self, ids, length: int = None) -> np.ndarray:
        return np.stack([self.tokens(int(i), length) for i in ids]) if len(ids) else np.zeros((0, length or self.cfg.retro_gpt_chunk_length), np.int64)
