"""``BlendedMegatronDatasetBuilder`` / ``BlendedMegatronDatasetConfig`` (reference
``core/datasets/blended_megatron_dataset_builder.py`` and ``blended_megatron_dataset_config.py``).

The builder turns (dataset class, per-split sample counts, rank predicate, config) into ``(train, valid, test)``:
a single corpus split by ``config.split``, a weighted blend of corpora, or independent blends per split
(``blend_per_split``).  ``is_built_on_rank()`` decides which ranks materialise datasets at all (e.g. only TP
rank 0 on the first / last pipeline stage); with ``torch.distributed`` initialised rank 0 builds first so that
the index caches are written once and every other rank memory-maps them.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple, Type

import numpy as np
import torch

from .gpt_dataset import BlendedDataset, GPTDataset, parse_split, split_document_ranges
from .indexed_dataset import IndexedDataset
from .mock import MockGPTDataset

Blend = Tuple[List[str], Optional[List[float]]]


@dataclass
class BlendedMegatronDatasetConfig:
    random_seed: int
    sequence_length: int
    blend: Optional[Blend] = None
    blend_per_split: Optional[List[Optional[Blend]]] = None
    split: Optional[str] = None
    path_to_cache: Optional[str] = None
    mmap_bin_files: bool = True
    mock: bool = False
    tokenizer: Optional[object] = None
    split_matrix: Optional[List[Optional[Tuple[float, float]]]] = field(init=False, default=None)

    def __post_init__(self):
        if self.blend_per_split is not None and any(b is not None for b in self.blend_per_split):
            if self.blend is not None:
                raise ValueError("blend and blend_per_split are incompatible")
            if self.split is not None:
                raise ValueError("split and blend_per_split are incompatible")
            if len(self.blend_per_split) != 3:
                raise ValueError("blend_per_split needs (train, valid, test)")
        elif self.blend is None:
            if not self.mock:
                self.mock = True              # nothing to read: synthetic data, as the reference does
            self.split = self.split or "1,1,1"
        if self.split is not None:
            w = parse_split(self.split)
            edges = np.concatenate([[0.0], np.cumsum(w)])
            self.split_matrix = [None if w[i] == 0 else (float(edges[i]), float(edges[i + 1])) for i in range(3)]


@dataclass
class GPTDatasetConfig(BlendedMegatronDatasetConfig):
    reset_position_ids: bool = False
    reset_attention_mask: bool = False
    eod_mask_loss: bool = False
    create_attention_mask: bool = True
    add_extra_token_to_sequence: bool = True


def get_blend_from_list(items: Optional[Sequence[str]]) -> Optional[Blend]:
    """``[w0, p0, w1, p1, ...]`` or ``[p0, p1, ...]`` -> ``(prefixes, weights | None)``."""
    if not items:
        return None
    items = list(items)
    if len(items) % 2 == 0:
        try:
            return [items[i + 1] for i in range(0, len(items), 2)], [float(items[i]) for i in range(0, len(items), 2)]
        except ValueError:
            pass
    return items, None


class BlendedMegatronDatasetBuilder:
    def __init__(self, cls: Type, sizes: Sequence[Optional[int]], is_built_on_rank: Callable[[], bool],
                 config: BlendedMegatronDatasetConfig):
        self.cls, self.sizes, self.is_built_on_rank, self.config = cls, list(sizes), is_built_on_rank, config
        if len(self.sizes) != 3:
            raise ValueError("sizes must be (train, valid, test)")

    # -- public -------------------------------------------------------------------
    def build(self) -> List[Optional[torch.utils.data.Dataset]]:
        dist = torch.distributed
        if not (dist.is_available() and dist.is_initialized()):
            return self._build() if self.is_built_on_rank() else [None, None, None]
        out = [None, None, None]
        if dist.get_rank() == 0 and self.is_built_on_rank():
            out = self._build()                       # writes the index caches
        dist.barrier()
        if dist.get_rank() != 0 and self.is_built_on_rank():
            out = self._build()
        return out

    # -- one dataset ---------------------------------------------------------------
    def _one(self, prefix: str, split_idx: int, n: Optional[int], frac: Optional[Tuple[float, float]]):
        c = self.config
        name = ("train", "valid", "test")[split_idx]
        if issubclass(self.cls, GPTDataset):
            ds = IndexedDataset(prefix, mmap=c.mmap_bin_files)
            ndoc = len(ds.document_indices) - 1
            lo, hi = (0, ndoc) if frac is None else (int(round(frac[0] * ndoc)), int(round(frac[1] * ndoc)))
            if hi <= lo:
                return None
            eod = c.tokenizer.eod if c.tokenizer is not None else 0
            n_eff = n if n is not None else max(1, int(ds.sequence_lengths[ds.document_indices[lo]:ds.document_indices[hi]].sum()
                                                       // c.sequence_length))
            return self.cls(ds, (lo, hi), n_eff, c.sequence_length, c.random_seed, name, c.path_to_cache, eod,
                            getattr(c, "reset_position_ids", False), getattr(c, "reset_attention_mask", False),
                            getattr(c, "eod_mask_loss", False), getattr(c, "create_attention_mask", False),
                            1 if getattr(c, "add_extra_token_to_sequence", True) else 0)
        return self.cls(prefix, split_idx, n, frac, c)       # user dataset classes: (prefix, split, n, doc range, config)

    def _mix(self, blend: Blend, split_idx: int, n: Optional[int], frac):
        prefixes, weights = blend
        if len(prefixes) == 1:
            return self._one(prefixes[0], split_idx, n, frac)
        if weights is None or n is None:
            parts = [self._one(p, split_idx, None, frac) for p in prefixes]
            parts = [p for p in parts if p is not None]
            return BlendedDataset(parts, [1.0] * len(parts), None) if parts else None
        w = np.asarray(weights, dtype=np.float64)
        w = w / w.sum()
        parts, ws = [], []
        for p, wi in zip(prefixes, w):
            d = self._one(p, split_idx, max(1, int(np.ceil(n * wi * 1.005))), frac)
            if d is not None:
                parts.append(d)
                ws.append(float(wi))
        return BlendedDataset(parts, ws, int(n)) if parts else None

    def _build(self) -> List[Optional[torch.utils.data.Dataset]]:
        c = self.config
        if c.mock:
            vocab = c.tokenizer.vocab_size if c.tokenizer is not None else 50257
            return [None if (n is None or n <= 0) else MockGPTDataset(n, c.sequence_length, vocab, c.random_seed + i)
                    for i, n in enumerate(self.sizes)]
        out: List[Optional[torch.utils.data.Dataset]] = []
        for i, n in enumerate(self.sizes):
            if c.blend_per_split is not None and any(b is not None for b in c.blend_per_split):
                b = c.blend_per_split[i]
                out.append(None if b is None or (n is not None and n <= 0) else self._mix(b, i, n, None))
            else:
                frac = c.split_matrix[i]
                out.append(None if frac is None or (n is not None and n <= 0) else self._mix(c.blend, i, n, frac))
        return out


# -- names of the reference's ``datasets/utils.py`` and ``blended_megatron_dataset_config.py`` ------------------------------------------
from enum import Enum as _Enum  # noqa: E402

from .gpt_dataset import convert_split_vector_to_split_matrix, parse_and_normalize_split  # noqa: E402,F401


class Split(_Enum):
    train = 0
    valid = 1
    test = 2


def normalize(weights):
    """Weights scaled to sum to one."""
    total = float(sum(weights))
    return [w / total for w in weights]


def compile_helpers():
    """Build the C++ index helpers (sample / shuffle / blending index construction).  They are compiled ahead of time by
    ``megatronapp_b200/build.py`` (``data/_helpers.so``); this (re)builds them when the shared object is missing."""
    import os

    here = os.path.dirname(os.path.abspath(__file__))
    if not os.path.exists(os.path.join(here, "_helpers.so")):
        from ..build import build_all

        build_all()
