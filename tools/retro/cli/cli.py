"""Inspect a Retro project (reference ``tools/retro/cli/cli.py``): the chunk database, the pre-training samples and the
neighbours retrieved for them, as token ids or as text.

    >>> from tools.retro.cli import retro
    >>> retro.init("work/retro")                       # project written by tools/retro/preprocess_data.py
    >>> retro.get_db_num_chunks()
    >>> retro.get_db_chunk_text(0)
    >>> retro.print_neighbor_texts(sample_id=0, chunk_id=1)
"""
from __future__ import annotations

import os
import sys
from typing import List, Optional, Tuple

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))


def shorten_str(s: str, n: int) -> str:
    s = "\\n".join(s.splitlines())
    return s if len(s) <= n else f"{s[:n // 2]} ... {s[-n // 2:]}"


class retro:
    cfg = None
    db = None
    tokenizer = None
    _pt = {}

    @classmethod
    def init(cls, project_dir: str, tokenizer=None) -> None:
        """``tokenizer``: anything with ``detokenize(ids)``; default: GPT-2 BPE from ``<project>/tokenizer/{vocab.json,merges.txt}``
        when present, else ids are printed as numbers."""
        from megatronapp_b200.data.retro import ChunkDB, RetroPreprocessingConfig

        cls.cfg = RetroPreprocessingConfig.load(project_dir)
        cls.db = ChunkDB(cls.cfg)
        cls._pt = {}
        cls.tokenizer = tokenizer
        tdir = os.path.join(project_dir, "tokenizer")
        if tokenizer is None and os.path.exists(os.path.join(tdir, "vocab.json")):
            from megatronapp_b200.data.tokenizer import GPT2BPETokenizer

            cls.tokenizer = GPT2BPETokenizer(os.path.join(tdir, "vocab.json"), os.path.join(tdir, "merges.txt"))

    # ---- text ------------------------------------------------------------------------------------------------------
    @classmethod
    def gpt_to_text(cls, token_ids) -> str:
        ids = [int(t) for t in np.asarray(token_ids).ravel()]
        return cls.tokenizer.detokenize(ids) if cls.tokenizer is not None else " ".join(map(str, ids))

    # ---- chunk database ----------------------------------------------------------------------------------------------
    @classmethod
    def get_db_num_indexed_datasets(cls) -> int:
        return len(cls.db.meta["prefixes"])

    @classmethod
    def get_db_indexed_dataset_infos(cls) -> List[Tuple[int, str]]:
        counts = np.bincount(np.asarray(cls.db.chunks[:, 0]), minlength=len(cls.db.meta["prefixes"]))
        return [(int(c), p) for c, p in zip(counts, cls.db.meta["prefixes"])]

    @classmethod
    def get_db_num_chunks(cls) -> int:
        return len(cls.db)

    @classmethod
    def get_db_chunk_gpt(cls, idx: int) -> List[int]:
        return cls.db.tokens(idx).tolist()

    @classmethod
    def get_db_chunk_text(cls, idx: int) -> str:
        return cls.gpt_to_text(cls.db.tokens(idx))

    @classmethod
    def get_db_chunk_and_continuation_text(cls, idx: int) -> List[str]:
        L = cls.cfg.retro_gpt_chunk_length
        both = cls.db.tokens(idx, cls.cfg.retro_retrieved_length)
        return [cls.gpt_to_text(both[:L]), cls.gpt_to_text(both[L:])]

    # ---- pre-training samples + neighbours -----------------------------------------------------------------------------
    @classmethod
    def get_pt_dataset(cls, data_key: str = "train"):
        if data_key not in cls._pt:
            from megatronapp_b200.data.retro import RetroDataset
            from megatronapp_b200.data.retro.query import build_train_dataset

            cls._pt[data_key] = RetroDataset(cls.cfg, build_train_dataset(cls.cfg), cls.db, data_key)
        return cls._pt[data_key]

    @classmethod
    def get_pt_num_samples_and_chunks(cls, data_key: str = "train") -> Tuple[int, int]:
        ds = cls.get_pt_dataset(data_key)
        return len(ds), len(ds) * cls.cfg.chunks_per_sample

    @classmethod
    def get_pt_num_samples(cls, data_key: str = "train") -> int:
        return cls.get_pt_num_samples_and_chunks(data_key)[0]

    @classmethod
    def get_pt_num_chunks(cls, data_key: str = "train") -> int:
        return cls.get_pt_num_samples_and_chunks(data_key)[1]

    @classmethod
    def get_pt_sample(cls, data_key: str, idx: int) -> dict:
        return cls.get_pt_dataset(data_key)[idx]

    @classmethod
    def get_neighbor_tokens(cls, sample_id: int, chunk_id: int, data_key: str = "train") -> Optional[dict]:
        ds = cls.get_pt_dataset(data_key)
        s = ds[sample_id]
        L = cls.cfg.retro_gpt_chunk_length
        return {"chunk_tokens": s["tokens"][chunk_id * L:(chunk_id + 1) * L].tolist(),
                "neighbor_ids": np.asarray(ds.neighbors[sample_id, chunk_id]).tolist(),
                "neighbor_tokens": s["neighbor_tokens"][chunk_id].tolist()}

    @classmethod
    def print_neighbor_texts(cls, sample_id: int, chunk_id: int, data_key: str = "train") -> None:
        t = cls.get_neighbor_tokens(sample_id, chunk_id, data_key)
        print(f"PRETRAINING CHUNK  (sample {sample_id}, chunk {chunk_id}):")
        print("  " + shorten_str(cls.gpt_to_text(t["chunk_tokens"]), 150))
        print("NEIGHBOR_CHUNKS (chunk + continuation):")
        for nid, toks in zip(t["neighbor_ids"], t["neighbor_tokens"]):
            print(f"  [{nid}] " + (shorten_str(cls.gpt_to_text(toks), 150) if nid >= 0 else "<none>"))

    @classmethod
    def print_usage(cls) -> None:
        print(f"retro project {cls.cfg.retro_project_dir}: {cls.get_db_num_chunks()} chunks of {cls.cfg.retro_gpt_chunk_length} "
              f"tokens from {cls.get_db_num_indexed_datasets()} dataset(s), index {cls.cfg.retro_index_str}, "
              f"{cls.cfg.retro_query_num_neighbors_save} neighbours x {cls.cfg.retro_retrieved_length} tokens per chunk")
        print("  retro.get_db_num_chunks() / get_db_chunk_gpt(i) / get_db_chunk_text(i) / get_db_chunk_and_continuation_text(i)")
        print("  retro.get_pt_num_samples() / get_pt_sample('train', i) / get_neighbor_tokens(sample, chunk) / print_neighbor_texts(sample, chunk)")
