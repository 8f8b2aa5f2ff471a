"""Masked-LM datasets on top of the sentence-level ``IndexedDataset`` (one sequence = one sentence,
documents = groups of sentences): BERT (sentence pairs + NSP/SOP, 80/10/10 masking) and T5
(span corruption with sentinel tokens).

Parity: reference ``core/datasets/masked_dataset.py`` (``MaskedWordPieceDataset`` sample mapping
through the C++ ``build_mapping`` helper, n-gram whole-word masking with the 80% [MASK] / 10%
random / 10% keep rule), ``bert_dataset.py`` (``BERTMaskedWordPieceDataset``: A/B segments,
random-next swap, [CLS]/[SEP], token types, padding mask) and ``t5_dataset.py``
(``T5MaskedWordPieceDataset``: sentinel replacement, encoder / decoder / cross masks).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .gpt_dataset import _helpers
from .indexed_dataset import IndexedDataset


@dataclass
class MaskedWordPieceDatasetConfig:
    sequence_length: int = 512
    masking_probability: float = 0.15
    short_sequence_probability: float = 0.1
    masking_max_ngram: int = 3
    masking_do_full_word: bool = True
    masking_do_permutation: bool = False
    masking_use_longer_ngrams: bool = False
    masking_use_geometric_distribution: bool = False
    random_seed: int = 1234
    classification_head: bool = True          # BERT NSP / SOP
    sequence_length_decoder: int = 128        # T5


class MaskedWordPieceDataset(torch.utils.data.Dataset):
    """Shared machinery: sample mapping over sentence spans + n-gram masking."""

    min_sentences = 1

    def __init__(self, indexed: IndexedDataset, num_samples: int, config: MaskedWordPieceDatasetConfig, tokenizer,
                 doc_range: Optional[Tuple[int, int]] = None, num_epochs: Optional[int] = None):
        self.ds, self.cfg, self.tok = indexed, config, tokenizer
        docs = np.asarray(indexed.document_indices, dtype=np.int64)
        if doc_range is not None:
            docs = docs[doc_range[0]:doc_range[1] + 1]
        sizes = np.asarray(indexed.sequence_lengths, dtype=np.int32)
        self.sample_index = _helpers().build_mapping(
            docs, sizes, int(num_epochs or 1000), int(num_samples), self._max_tokens(), config.short_sequence_probability, config.random_seed, False,
            self.min_sentences)
        self.vocab_ids = np.arange(getattr(tokenizer, "vocab_size", 30522))
        self.cls, self.sep = tokenizer.cls, tokenizer.sep
        self.mask_id, self.pad = tokenizer.mask, tokenizer.pad
        inv = getattr(tokenizer, "inv_vocab", None)
        self._is_continuation = (lambda t: inv[int(t)].startswith("##")) if inv is not None else (lambda t: False)

    def _max_tokens(self) -> int:
        return self.cfg.sequence_length - 3

    def __len__(self):
        return int(self.sample_index.shape[0])

    def _sentences(self, idx) -> Tuple[List[np.ndarray], int]:
        first, last, target = (int(x) for x in self.sample_index[idx])
        return [np.asarray(self.ds[i]) for i in range(first, last)], target

    def _mask(self, tokens: np.ndarray, rng: np.random.RandomState, special: Sequence[int]):
        """Returns (masked tokens, positions, original labels, span boundaries)."""
        cfg = self.cfg
        n = len(tokens)
        cand: List[List[int]] = []
        for i, t in enumerate(tokens):
            if int(t) in special:
                continue
            if cfg.masking_do_full_word and cand and self._is_continuation(t):
                cand[-1].append(i)
            else:
                cand.append([i])
        budget = max(1, int(round((n - sum(int(t) in special for t in tokens)) * cfg.masking_probability)))
        order = rng.permutation(len(cand))
        out = tokens.copy()
        covered = set()
        positions, labels, spans = [], [], []
        ngrams = np.arange(1, cfg.masking_max_ngram + 1)
        pvals = 1.0 / ngrams
        if cfg.masking_use_longer_ngrams:
            pvals = pvals[::-1]
        pvals = pvals / pvals.sum()
        for ci in order:
            if len(positions) >= budget:
                break
            if cfg.masking_use_geometric_distribution:
                k = min(int(rng.geometric(0.2)), cfg.masking_max_ngram)
            else:
                k = int(rng.choice(ngrams, p=pvals))
            idxs = [i for grp in cand[ci:ci + k] for i in grp]
            if not idxs or len(positions) + len(idxs) > budget or any(i in covered for i in idxs):
                continue
            covered.update(idxs)
            spans.append((idxs[0], idxs[-1] + 1))
            for i in idxs:
                r = rng.random_sample()
                if r < 0.8:
                    out[i] = self.mask_id
                elif r < 0.9:
                    out[i] = self.vocab_ids[rng.randint(0, len(self.vocab_ids))]
                positions.append(i)
                labels.append(int(tokens[i]))
        o = np.argsort(positions)
        return out, [positions[i] for i in o], [labels[i] for i in o], sorted(spans)


class BERTMaskedWordPieceDataset(MaskedWordPieceDataset):
    min_sentences = 2

    def __getitem__(self, idx) -> Dict[str, torch.Tensor]:
        cfg = self.cfg
        sents, target = self._sentences(idx)
        rng = np.random.RandomState(seed=(cfg.random_seed + idx) % 2 ** 32)
        if cfg.classification_head and len(sents) > 1:
            pivot = rng.randint(1, len(sents))
            a = np.concatenate(sents[:pivot])
            b = np.concatenate(sents[pivot:])
            is_random = int(rng.random_sample() < 0.5)
            if is_random:
                a, b = b, a
        else:
            a, b, is_random = np.concatenate(sents), np.zeros(0, dtype=np.int64), 0
        # truncate the longer segment from a random end until the pair fits
        budget = min(target, self._max_tokens())
        a, b = list(a), list(b)
        while len(a) + len(b) > budget:
            seg = a if len(a) > len(b) else b
            seg.pop(0) if rng.random_sample() < 0.5 else seg.pop()
        toks = [self.cls] + a + [self.sep] + (b + [self.sep] if b else [])
        types = [0] * (len(a) + 2) + [1] * (len(b) + 1 if b else 0)
        toks = np.asarray(toks, dtype=np.int64)
        masked, pos, lab, _ = self._mask(toks, rng, (self.cls, self.sep))
        L = cfg.sequence_length
        n = len(masked)
        text = np.full(L, self.pad, dtype=np.int64)
        text[:n] = masked
        tt = np.zeros(L, dtype=np.int64)
        tt[:n] = types
        labels = np.full(L, -1, dtype=np.int64)
        loss_mask = np.zeros(L, dtype=np.float32)
        labels[pos] = lab
        loss_mask[pos] = 1.0
        pad_mask = np.zeros(L, dtype=np.int64)
        pad_mask[:n] = 1
        return {"text": torch.from_numpy(text), "types": torch.from_numpy(tt),
                "labels": torch.from_numpy(np.where(labels < 0, 0, labels)), "is_random": torch.tensor(is_random),
                "loss_mask": torch.from_numpy(loss_mask), "padding_mask": torch.from_numpy(pad_mask),
                "truncated": torch.tensor(int(len(a) + len(b) >= budget))}


class T5MaskedWordPieceDataset(MaskedWordPieceDataset):
    """Span corruption: each masked span becomes one sentinel in the encoder input; the decoder
    target is ``<sentinel_k> span_k ...`` terminated by EOS."""

    def __init__(self, *a, sentinel_tokens: Optional[Sequence[int]] = None, **k):
        super().__init__(*a, **k)
        st = sentinel_tokens or getattr(self.tok, "additional_special_tokens_ids", None)
        if not st:
            v = getattr(self.tok, "vocab_size", 30522)
            st = list(range(v - 100, v))
        self.sentinels = list(st)
        self.bos = getattr(self.tok, "bos", None) or self.cls
        self.eos = getattr(self.tok, "eos", None) or self.sep

    def _max_tokens(self) -> int:
        return self.cfg.sequence_length - 2

    def __getitem__(self, idx) -> Dict[str, torch.Tensor]:
        cfg = self.cfg
        sents, target = self._sentences(idx)
        rng = np.random.RandomState(seed=(cfg.random_seed + idx) % 2 ** 32)
        toks = np.concatenate(sents)[:min(target, self._max_tokens())].astype(np.int64)
        _, _, _, spans = self._mask(toks, rng, ())
        enc, dec_in, dec_out = [], [self.bos], []
        cur = 0
        for k, (s, e) in enumerate(spans[:len(self.sentinels)]):
            sent = self.sentinels[k]
            enc.extend(toks[cur:s].tolist())
            enc.append(sent)
            dec_in.append(sent)
            dec_in.extend(toks[s:e].tolist())
            dec_out.append(sent)
            dec_out.extend(toks[s:e].tolist())
            cur = e
        enc.extend(toks[cur:].tolist())
        dec_out.append(self.eos)
        Le, Ld = cfg.sequence_length, cfg.sequence_length_decoder
        enc, dec_in, dec_out = enc[:Le], dec_in[:Ld], dec_out[:Ld]

        def pad(x, L):
            out = np.full(L, self.pad, dtype=np.int64)
            out[:len(x)] = x
            return out

        ne, nd = len(enc), len(dec_in)
        enc_valid = (np.arange(Le) < ne)
        dec_valid = (np.arange(Ld) < nd)
        loss_mask = (np.arange(Ld) < len(dec_out)).astype(np.float32)
        # True = attend (converted to Megatron's True = masked-out convention by the model wrapper)
        enc_mask = enc_valid[:, None] & enc_valid[None, :]
        dec_mask = (dec_valid[:, None] & dec_valid[None, :]) & np.tril(np.ones((Ld, Ld), dtype=bool))
        enc_dec_mask = dec_valid[:, None] & enc_valid[None, :]
        return {"text_enc": torch.from_numpy(pad(enc, Le)), "text_dec": torch.from_numpy(pad(dec_in, Ld)),
                "labels": torch.from_numpy(pad(dec_out, Ld)), "loss_mask": torch.from_numpy(loss_mask),
                "enc_mask": torch.from_numpy(enc_mask), "dec_mask": torch.from_numpy(dec_mask),
                "enc_dec_mask": torch.from_numpy(enc_dec_mask), "truncated": torch.tensor(int(len(toks) >= target))}
