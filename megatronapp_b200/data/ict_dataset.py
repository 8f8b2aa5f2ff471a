"""Inverse-cloze-task dataset (reference ``megatron/legacy/data/ict_dataset.py``): a *block* is a run of consecutive
sentences of one document that fits the context length; a sample picks one sentence of the block as the pseudo
query and (with probability ``1 - query_in_block_prob``) removes it from the block, which becomes the context."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .indexed_dataset import IndexedDataset


class ICTDataset(torch.utils.data.Dataset):
    def __init__(self, block_dataset: IndexedDataset, title_dataset: Optional[IndexedDataset], num_samples: Optional[int],
                 max_seq_length: int, query_in_block_prob: float, seed: int, cls_id: int, sep_id: int, pad_id: int,
                 use_titles: bool = True, use_one_sent_docs: bool = False):
        self.ds, self.titles, self.max_seq_length = block_dataset, title_dataset, max_seq_length
        self.query_in_block_prob, self.seed = query_in_block_prob, seed
        self.cls, self.sep, self.pad = cls_id, sep_id, pad_id
        self.use_titles = use_titles and title_dataset is not None
        self.blocks = self._build_blocks(use_one_sent_docs)              # rows: (first sentence, end sentence, doc, block id)
        self.num_samples = num_samples if num_samples is not None else len(self.blocks)

    def _build_blocks(self, one_sent_ok: bool) -> np.ndarray:
        di, lens = self.ds.document_indices, self.ds.sequence_lengths
        rows: List[tuple] = []
        for d in range(len(di) - 1):
            lo, hi = int(di[d]), int(di[d + 1])
            if hi - lo < (1 if one_sent_ok else 2):
                continue
            budget = self.max_seq_length - 3 - (len(self.titles[d]) if self.use_titles else 0)
            start, used = lo, 0
            for s in range(lo, hi):
                n = int(lens[s])
                if used + n > budget and s > start:
                    if s - start >= (1 if one_sent_ok else 2):
                        rows.append((start, s, d, len(rows)))
                    start, used = s, 0
                used += n
            if hi - start >= (1 if one_sent_ok else 2):
                rows.append((start, hi, d, len(rows)))
        return np.asarray(rows, dtype=np.int64).reshape(-1, 4)

    def __len__(self):
        return self.num_samples

    def _pad(self, ids: List[int]):
        ids = ids[: self.max_seq_length]
        n = len(ids)
        out = np.full(self.max_seq_length, self.pad, dtype=np.int64)
        out[:n] = ids
        mask = np.zeros(self.max_seq_length, dtype=np.int64)
        mask[:n] = 1
        return out, mask

    def get_block(self, start: int, end: int, doc: int):
        """Context tokens of a block by its sentence range (used when the evidence index is built)."""
        title = list(self.titles[doc]) if self.use_titles else None
        return self._context([list(self.ds[i]) for i in range(start, end)], title)

    def _context(self, sents: List[list], title: Optional[list]):
        flat = [t for s in sents for t in s]
        if title is not None:
            flat = flat[: self.max_seq_length - 3 - len(title)]
            return self._pad([self.cls] + title + [self.sep] + flat + [self.sep])
        return self._pad([self.cls] + flat[: self.max_seq_length - 2] + [self.sep])

    def __getitem__(self, idx):
        start, end, doc, block_id = (int(v) for v in self.blocks[idx % len(self.blocks)])
        rng = np.random.default_rng(self.seed + idx)
        sents = [list(self.ds[i]) for i in range(start, end)]
        q = int(rng.integers(0, len(sents)))
        query = sents[q] if rng.random() < self.query_in_block_prob or len(sents) == 1 else sents.pop(q)
        qt, qm = self._pad([self.cls] + list(query)[: self.max_seq_length - 2] + [self.sep])
        ct, cm = self._context(sents, list(self.titles[doc]) if self.use_titles else None)
        return {"query_tokens": qt, "query_pad_mask": qm, "context_tokens": ct, "context_pad_mask": cm,
                "block_data": np.asarray([start, end, doc, block_id], dtype=np.int64)}
