"""GPT pre-training dataset over an indexed corpus: document / sample / shuffle indices, caching,
train/valid/test splits and weighted blending.

Parity: reference ``core/datasets/gpt_dataset.py`` (``GPTDataset``: ``_build_document_sample_shuffle_indices``
with the separate last epoch rule, index cache files, EOD masking options), ``blended_dataset.py``
and ``blended_megatron_dataset_builder.py`` (``--split 969,30,1``, ``--data-path w1 p1 w2 p2``).
"""
from __future__ import annotations

import hashlib
import json
import os
from typing import List, Optional, Tuple

import numpy as np
import torch

from .indexed_dataset import IndexedDataset
from .mock import get_ltor_masks_and_position_ids


def _helpers():
    import importlib.util
    from pathlib import Path

    so = Path(__file__).resolve().parent / "_helpers.so"
    if not so.exists():
        from ..build import build_all

        build_all()
    spec = importlib.util.spec_from_file_location("_helpers", str(so))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def parse_split(split: str) -> List[float]:
    parts = [float(x) for x in split.replace("/", ",").split(",") if x != ""]
    while len(parts) < 3:
        parts.append(0.0)
    tot = sum(parts[:3])
    return [p / tot for p in parts[:3]]


def split_document_ranges(num_docs: int, split: List[float]) -> List[Tuple[int, int]]:
    bounds = [0]
    for s in split:
        bounds.append(bounds[-1] + int(round(s * num_docs)))
    diff = bounds[-1] - num_docs
    bounds = [bounds[0]] + [b - diff for b in bounds[1:]]
    return [(bounds[i], bounds[i + 1]) for i in range(3)]


class GPTDataset(torch.utils.data.Dataset):
    def __init__(self, indexed: IndexedDataset, doc_range: Tuple[int, int], num_samples: int, seq_length: int,
                 seed: int, name: str = "train", cache_dir: Optional[str] = None, eod_token: int = 0,
                 reset_position_ids=False, reset_attention_mask=False, eod_mask_loss=False,
                 create_attention_mask=False, add_extra_token: int = 1):
        self.ds, self.seq_length, self.name = indexed, seq_length, name
        self.eod, self.add_extra = eod_token, add_extra_token
        self.reset_position_ids, self.reset_attention_mask = reset_position_ids, reset_attention_mask
        self.eod_mask_loss, self.create_attention_mask = eod_mask_loss, create_attention_mask
        self.documents = np.arange(doc_range[0], doc_range[1], dtype=np.int32)
        key = hashlib.md5(json.dumps([indexed.path_prefix, list(map(int, doc_range)), num_samples, seq_length, seed,
                                      name]).encode()).hexdigest()
        cache = os.path.join(cache_dir, key) if cache_dir else None
        if cache and all(os.path.exists(f"{cache}-{n}.npy") for n in ("doc", "sample", "shuffle")):
            self.doc_idx = np.load(f"{cache}-doc.npy", mmap_mode="r")
            self.sample_idx = np.load(f"{cache}-sample.npy", mmap_mode="r")
            self.shuffle_idx = np.load(f"{cache}-shuffle.npy", mmap_mode="r")
        else:
            self.doc_idx, self.sample_idx, self.shuffle_idx = self._build(num_samples, seed)
            if cache and (not torch.distributed.is_initialized() or torch.distributed.get_rank() == 0):
                os.makedirs(cache_dir, exist_ok=True)
                np.save(f"{cache}-doc.npy", self.doc_idx)
                np.save(f"{cache}-sample.npy", self.sample_idx)
                np.save(f"{cache}-shuffle.npy", self.shuffle_idx)

    def _build(self, num_samples: int, seed: int):
        sizes = np.asarray(self.ds.sequence_lengths)
        tokens_per_epoch = int(sizes[self.documents].sum())
        if tokens_per_epoch <= self.seq_length:
            raise ValueError("dataset split has fewer tokens than one sequence")
        num_epochs = 1
        while (num_epochs * tokens_per_epoch - self.add_extra) // self.seq_length < num_samples:
            num_epochs += 1
        rng = np.random.RandomState(seed)
        # the last epoch is shuffled separately when it contributes < 80 % of an epoch of samples
        separate_last = False
        if num_epochs > 1:
            from_prev = ((num_epochs - 1) * tokens_per_epoch - self.add_extra) // self.seq_length
            last_needed = num_samples - from_prev
            per_epoch = (tokens_per_epoch - self.add_extra) // self.seq_length
            separate_last = last_needed < int(0.80 * per_epoch)
        if separate_last:
            head = np.tile(self.documents, num_epochs - 1)
            rng.shuffle(head)
            tail = self.documents.copy()
            rng.shuffle(tail)
            doc_idx = np.concatenate([head, tail]).astype(np.int32)
        else:
            doc_idx = np.tile(self.documents, num_epochs).astype(np.int32)
            rng.shuffle(doc_idx)
        h = _helpers()
        sample_idx = h.build_sample_idx_int32(np.ascontiguousarray(sizes, dtype=np.int32), doc_idx, self.seq_length,
                                              num_epochs, tokens_per_epoch, True, self.add_extra)
        n = sample_idx.shape[0] - 1
        shuffle_idx = np.arange(n, dtype=np.int64)
        rng.shuffle(shuffle_idx)
        return doc_idx, sample_idx, shuffle_idx

    def __len__(self):
        return self.sample_idx.shape[0] - 1

    def _tokens(self, idx: int) -> np.ndarray:
        idx = int(self.shuffle_idx[idx % len(self)])
        d0, o0 = self.sample_idx[idx]
        d1, o1 = self.sample_idx[idx + 1]
        if d0 == d1:
            return np.asarray(self.ds.get(int(self.doc_idx[d0]), int(o0), int(o1 - o0 + self.add_extra)), dtype=np.int64)
        parts = [self.ds.get(int(self.doc_idx[d0]), int(o0))]
        for d in range(d0 + 1, d1):
            parts.append(self.ds.get(int(self.doc_idx[d])))
        parts.append(self.ds.get(int(self.doc_idx[d1]), 0, int(o1 + self.add_extra)))
        return np.concatenate(parts).astype(np.int64)

    def __getitem__(self, idx):
        text = torch.from_numpy(self._tokens(idx))
        tokens, labels = text[:-1].contiguous(), text[1:].contiguous()
        att, loss_mask, pos = get_ltor_masks_and_position_ids(tokens.unsqueeze(0), self.eod, self.reset_position_ids,
                                                              self.reset_attention_mask, self.eod_mask_loss,
                                                              self.create_attention_mask)
        out = {"tokens": tokens, "labels": labels, "loss_mask": loss_mask[0], "position_ids": pos[0]}
        if att is not None:
            out["attention_mask"] = att[0]
        return out


class BlendedDataset(torch.utils.data.Dataset):
    def __init__(self, datasets: List[torch.utils.data.Dataset], weights: List[float], size: Optional[int]):
        self.datasets = datasets
        h = _helpers()
        if size is None:
            sizes = np.array([len(d) for d in datasets], dtype=np.int64)
            size = int(sizes.sum())
            self.dataset_index = np.zeros(size, dtype=np.int16)
            self.dataset_sample_index = np.zeros(size, dtype=np.int64)
            h.build_exhaustive_blending_indices(self.dataset_index, self.dataset_sample_index, sizes, len(datasets))
        else:
            w = np.array(weights, dtype=np.float64)
            w /= w.sum()
            self.dataset_index = np.zeros(size, dtype=np.int16)
            self.dataset_sample_index = np.zeros(size, dtype=np.int64)
            h.build_blending_indices(self.dataset_index, self.dataset_sample_index, w, len(datasets), size, False)
        self.size = size

    def __len__(self):
        return self.size

    def __getitem__(self, idx):
        d, s = int(self.dataset_index[idx]), int(self.dataset_sample_index[idx])
        return self.datasets[d][s % len(self.datasets[d])]


def parse_blend(data_path: List[str]):
    if len(data_path) == 1:
        return [data_path[0]], [1.0]
    if len(data_path) % 2 != 0:
        raise ValueError("--data-path expects either one prefix or weight/prefix pairs")
    return [data_path[i + 1] for i in range(0, len(data_path), 2)], [float(data_path[i]) for i in range(0, len(data_path), 2)]


def blend_and_blend_per_split(args):
    """(blend, per-split blends) from the data flags, as the reference's ``get_blend_and_blend_per_split``: ``--data-path`` or a
    file holding it (``--data-args-path``) give ONE blend cut by ``--split``; ``--train-data-path / --valid-data-path /
    --test-data-path`` or a JSON ``{"train": [...], "valid": [...], "test": [...]}`` (``--per-split-data-args-path``) give each
    split its own blend, used whole."""
    import json

    blend = getattr(args, "data_path", None)
    if getattr(args, "data_args_path", None):
        with open(args.data_args_path) as f:
            blend = f.read().split()
    per_split = [getattr(args, f"{n}_data_path", None) for n in ("train", "valid", "test")]
    if getattr(args, "per_split_data_args_path", None):
        with open(args.per_split_data_args_path) as f:
            cfg = json.load(f)
        per_split = [cfg.get(n) for n in ("train", "valid", "test")]
    if any(per_split):
        if blend:
            raise ValueError("give either --data-path (one blend, cut by --split) or the per-split data paths, not both")
        return None, per_split
    return blend, None


def build_gpt_datasets(args, train_val_test_num_samples):
    eod = 0
    tok = None
    try:
        from ..training.global_vars import get_tokenizer

        tok = get_tokenizer()
        eod = tok.eod if tok is not None else 0
    except Exception:
        pass

    def dataset(ds, doc_range, n, name):
        return GPTDataset(ds, doc_range, n, args.seq_length, args.seed, name, args.data_cache_path, eod, args.reset_position_ids,
                          args.reset_attention_mask, args.eod_mask_loss, args.create_attention_mask_in_dataloader)

    def blended(parts, weights, n):
        parts_w = [(d, w) for d, w in zip(parts, weights) if d is not None]
        if not parts_w:
            return None
        if len(parts_w) == 1:
            return parts_w[0][0]
        return BlendedDataset([d for d, _ in parts_w], [w for _, w in parts_w], int(n))

    blend, blend_per_split = blend_and_blend_per_split(args)
    names = ("train", "valid", "test")
    if blend_per_split is not None:
        out = []
        for name, paths, n in zip(names, blend_per_split, train_val_test_num_samples):
            if not paths or n <= 0:
                out.append(None)
                continue
            prefixes, weights = parse_blend(paths)
            parts = []
            for prefix, w in zip(prefixes, weights):
                ds = IndexedDataset(prefix, mmap=args.mmap_bin_files)
                ndoc = len(ds.document_indices) - 1
                parts.append(dataset(ds, (0, ndoc), max(1, int(np.ceil(n * w / sum(weights) * 1.005))), name) if ndoc > 0 else None)
            out.append(blended(parts, weights, n))
        return tuple(out)
    prefixes, weights = parse_blend(blend)
    split = parse_split(args.split or "969,30,1")
    per_split: List[List[GPTDataset]] = [[], [], []]
    for prefix, w in zip(prefixes, weights):
        ds = IndexedDataset(prefix, mmap=args.mmap_bin_files)
        ranges = split_document_ranges(len(ds.document_indices) - 1, split)
        for i, (name, n) in enumerate(zip(names, train_val_test_num_samples)):
            lo, hi = ranges[i]
            if hi - lo <= 0 or n <= 0:
                per_split[i].append(None)
                continue
            per_split[i].append(dataset(ds, (lo, hi), max(1, int(np.ceil(n * w / sum(weights) * 1.005))), name))
    return tuple(blended(per_split[i], weights, n) for i, n in enumerate(train_val_test_num_samples))


def __getattr__(name):      # names the reference exposes from ``core/datasets/gpt_dataset.py`` that live in sibling modules here
    if name in ("GPTDatasetConfig", "BlendedMegatronDatasetConfig"):
        from . import blended_builder

        return getattr(blended_builder, name)
    if name == "MockGPTDataset":
        from .mock import MockGPTDataset

        return MockGPTDataset
    raise AttributeError(name)


def parse_and_normalize_split(split: str) -> List[float]:
    """``"949,50,1"`` -> normalised train / valid / test fractions (reference ``blended_megatron_dataset_config``)."""
    return parse_split(split)


def convert_split_vector_to_split_matrix(vector_a: List[float], vector_b: Optional[List[float]] = None):
    """Per split the ``(begin, end)`` fraction of the document range it covers (``None`` for an empty split); with a second
    vector, the overlap of the two bookings (used to carve e.g. a retrieval split out of the same corpus)."""
    def matrix(v):
        out, lo = [], 0.0
        for f in v:
            out.append((lo, lo + f) if f > 0 else None)
            lo += f
        return out

    a = matrix(vector_a)
    if vector_b is None:
        return a
    b = matrix(vector_b)
    res = []
    for x, y in zip(a, b):
        if x is None or y is None:
            res.append(None)
            continue
        lo, hi = max(x[0], y[0]), min(x[1], y[1])
        res.append((lo, hi) if hi > lo else None)
    return res
