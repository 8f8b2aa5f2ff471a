"""Preprocessing configuration (reference ``core/datasets/retro/config/config.py`` ``RetroPreprocessingConfig``)."""
from __future__ import annotations

import json
import os
from dataclasses import asdict, dataclass, field
from typing import List, Optional


@dataclass
class RetroPreprocessingConfig:
    retro_project_dir: str
    retro_gpt_data_path: List[str] = field(default_factory=list)     # [weight] prefix ... (the training blend)
    retro_gpt_seq_length: int = 2048
    retro_gpt_chunk_length: int = 64
    retro_gpt_split: str = "969,30,1"
    retro_gpt_seed: int = 1234
    retro_gpt_train_samples: int = 0
    retro_gpt_eod: int = 0
    retro_vocab_size: int = 50304
    retro_index_str: str = "IVF256,PQ16"       # faiss-style factory string: "Flat" | "IVF<nlist>,PQ<m>" | "IVF<nlist>,Flat"
    retro_index_ntrain: int = 100_000
    retro_index_train_load_fraction: float = 1.0
    retro_index_add_load_fraction: float = 1.0
    retro_query_nprobe: int = 16
    retro_query_num_neighbors_query: int = 20  # searched
    retro_query_num_neighbors_save: int = 2    # kept after removing same-document hits
    retro_retrieved_length: int = 128          # neighbour chunk + continuation
    retro_block_size: int = 10_000             # samples per query block (resume granularity)
    retro_embed_dim: int = 256
    retro_embedder: str = "hash"               # "hash" | "bert:<checkpoint dir>" (tools/bert_embedding)
    retro_doc_block_size: int = 100_000

    def path(self, *parts: str) -> str:
        return os.path.join(self.retro_project_dir, *parts)

    def save(self) -> None:
        os.makedirs(self.retro_project_dir, exist_ok=True)
        with open(self.path("config.json"), "w") as f:
            json.dump(asdict(self), f, indent=1)

    @classmethod
    def load(cls, project_dir: str) -> "RetroPreprocessingConfig":
        with open(os.path.join(project_dir, "config.json")) as f:
            d = json.load(f)
        d["retro_project_dir"] = project_dir
        return cls(**d)

    @property
    def chunks_per_sample(self) -> int:
        return self.retro_gpt_seq_length // self.retro_gpt_chunk_length
