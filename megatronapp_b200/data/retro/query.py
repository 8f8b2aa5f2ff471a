"""Neighbour query + training dataset (reference ``core/datasets/retro/query/{query,retro_dataset,gpt_chunk_dataset}.py``).

For every pre-training sample (the *same* GPTDataset order the trainer will see: same blend, split, seed, sequence length)
and every ``chunk_length`` window of it, search the index for ``num_neighbors_query`` chunks, drop hits that come from a
document the sample itself was cut from (no leaking the continuation the model has to predict), keep
``num_neighbors_save``.  Results are written block by block into one ``.npy`` memory map so an interrupted run resumes."""
from __future__ import annotations

import json
import os
from typing import Callable, Optional

import numpy as np
import torch

from .config import RetroPreprocessingConfig
from .db import ChunkDB


def hash_embedder(vocab_size: int, dim: int, seed: int = 0) -> Callable[[np.ndarray], torch.Tensor]:
    """Deterministic, corpus independent bag-of-tokens embedding (random projection of token counts); the reference uses
    mean-pooled BERT (``config/bert_embedders.py``) -- pass ``retro_embedder='bert:<ckpt>'`` for that."""
    g = torch.Generator().manual_seed(seed)
    table = torch.randn(vocab_size, dim, generator=g)
    dev = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    table = table.to(dev)

    def embed(tokens: np.ndarray) -> torch.Tensor:
        t = torch.from_numpy(np.ascontiguousarray(tokens)).to(dev).long().clamp_(0, vocab_size - 1)
        return torch.nn.functional.normalize(table[t].mean(1), dim=-1)

    return embed


def get_embedder(cfg: RetroPreprocessingConfig, bert_model=None):
    """``bert_model``: a BERT encoder (``tools/bert_embedding.BertEmbedder`` mean-pools its last layer) -- the chunk tokens
    must then be in that model's vocabulary; default: the hashing embedder."""
    if bert_model is not None:
        import sys

        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
        from tools.bert_embedding.embed import BertEmbedder

        be = BertEmbedder(bert_model, max_seq_length=cfg.retro_gpt_chunk_length)
        return lambda toks: torch.nn.functional.normalize(be.embed_tokens(torch.from_numpy(np.ascontiguousarray(toks))).float(), dim=-1)
    return hash_embedder(cfg.retro_vocab_size, cfg.retro_embed_dim, cfg.retro_gpt_seed)


def build_train_dataset(cfg: RetroPreprocessingConfig, num_samples: Optional[int] = None, split_index: int = 0):
    """The GPT dataset exactly as ``pretrain_retro.py`` builds it from the same arguments."""
    from types import SimpleNamespace

    from ..gpt_dataset import build_gpt_datasets

    n = num_samples or cfg.retro_gpt_train_samples
    a = SimpleNamespace(data_path=cfg.retro_gpt_data_path, split=cfg.retro_gpt_split, seq_length=cfg.retro_gpt_seq_length,
                        seed=cfg.retro_gpt_seed, data_cache_path=None, reset_position_ids=False, reset_attention_mask=False,
                        eod_mask_loss=False, create_attention_mask_in_dataloader=False, mmap_bin_files=True)
    sizes = [0, 0, 0]
    sizes[split_index] = n
    return build_gpt_datasets(a, sizes)[split_index]


def _sample_doc_keys(ds, idx: int):
    """(dataset index, document) pairs a GPT sample was cut from"""
    if hasattr(ds, "datasets"):                                      # BlendedDataset
        di, si = int(ds.dataset_index[idx]), int(ds.dataset_sample_index[idx])
        return {(di, d) for _, d in _sample_doc_keys(ds.datasets[di], si)}
    i = int(ds.shuffle_idx[idx % len(ds)])
    d0, d1 = int(ds.sample_idx[i][0]), int(ds.sample_idx[i + 1][0])
    return {(0, int(ds.doc_idx[d])) for d in range(d0, d1 + 1)}


def query_neighbors(cfg: RetroPreprocessingConfig, db: ChunkDB, index, dataset, name: str = "train",
                    embed: Optional[Callable] = None) -> np.ndarray:
    embed = embed or get_embedder(cfg)
    n, C, K = len(dataset), cfg.chunks_per_sample, cfg.retro_query_num_neighbors_save
    os.makedirs(cfg.path("query"), exist_ok=True)
    path, prog = cfg.path("query", f"{name}_neighbors.npy"), cfg.path("query", f"{name}_progress.json")
    done = 0
    if os.path.exists(path) and os.path.exists(prog):
        st = json.load(open(prog))
        if st.get("n") == n and st.get("C") == C and st.get("K") == K:
            done = int(st["done"])
    out = np.lib.format.open_memmap(path, mode="r+" if done else "w+", dtype=np.int64, shape=(n, C, K))
    L = cfg.retro_gpt_chunk_length
    for b0 in range(done, n, cfg.retro_block_size):
        b1 = min(n, b0 + cfg.retro_block_size)
        toks = np.stack([np.asarray(dataset[i]["tokens"])[: C * L] for i in range(b0, b1)]).reshape(-1, L)
        _, ids = index.search(embed(toks), cfg.retro_query_num_neighbors_query, cfg.retro_query_nprobe)
        ids = ids.reshape(b1 - b0, C, -1)
        for i in range(b0, b1):
            own = _sample_doc_keys(dataset, i)
            for c in range(C):
                keep = [int(j) for j in ids[i - b0, c] if j >= 0 and db.doc_key(int(j)) not in own][:K]
                out[i, c, :len(keep)] = keep
                out[i, c, len(keep):] = -1
        out.flush()
        json.dump({"n": n, "C": C, "K": K, "done": b1}, open(prog, "w"))
    return out


class RetroDataset(torch.utils.data.Dataset):
    """GPT sample + ``neighbor_tokens [chunks, neighbours, retrieved_length]`` (missing neighbours are all-pad)."""

    def __init__(self, cfg: RetroPreprocessingConfig, dataset, db: ChunkDB, name: str = "train", pad: int = 0,
                 verify_neighbor_count: bool = True):
        self.cfg, self.ds, self.db, self.pad = cfg, dataset, db, pad
        self.neighbors = np.load(cfg.path("query", f"{name}_neighbors.npy"), mmap_mode="r")
        # reference retro_dataset.py:216 (--retro-no-verify-neighbor-count skips it: index with idx % table size instead)
        if verify_neighbor_count and self.neighbors.shape[0] != len(dataset):
            raise ValueError(f"neighbour table covers {self.neighbors.shape[0]} samples, dataset has {len(dataset)}: "
                             "query with the same data / split / seed / sequence length as training")

    def __len__(self):
        return len(self.ds)

    def __getitem__(self, idx):
        s = dict(self.ds[idx])
        nb = self.neighbors[(idx % len(self.ds)) % self.neighbors.shape[0]]
        R = self.cfg.retro_retrieved_length
        toks = np.full((*nb.shape, R), self.pad, dtype=np.int64)
        for c in range(nb.shape[0]):
            for k in range(nb.shape[1]):
                if nb[c, k] >= 0:
                    toks[c, k] = self.db.tokens(int(nb[c, k]), R, self.pad)
        s["neighbor_tokens"] = torch.from_numpy(toks)
        return s
