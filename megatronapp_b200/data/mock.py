"""Synthetic datasets (``--mock-data``): deterministic token streams of the requested shape.

Parity: reference ``core/datasets/gpt_dataset.py MockGPTDataset`` / ``pretrain_gpt.py:298-301``.
"""
from __future__ import annotations

import numpy as np
import torch


def get_ltor_masks_and_position_ids(data, eod_token, reset_position_ids=False, reset_attention_mask=False,
                                    eod_mask_loss=False, create_attention_mask=False):
    """Left-to-right masks and position ids (reference ``training/utils.py``)."""
    b, s = data.size()
    att = None
    if create_attention_mask:
        att_b = b if reset_attention_mask else 1
        att = torch.tril(torch.ones((att_b, s, s), device=data.device)).view(att_b, 1, s, s)
    loss_mask = torch.ones(data.size(), dtype=torch.float, device=data.device)
    if eod_mask_loss:
        loss_mask[data == eod_token] = 0.0
    position_ids = torch.arange(s, dtype=torch.long, device=data.device).unsqueeze(0).expand_as(data)
    if reset_position_ids:
        position_ids = position_ids.clone()
    if reset_position_ids or reset_attention_mask:
        for bi in range(b):
            eod_idx = position_ids[bi, data[bi] == eod_token]
            if reset_position_ids:
                eod_idx = eod_idx.clone()
            prev = 0
            for j in range(eod_idx.size(0)):
                i = eod_idx[j]
                if reset_attention_mask and att is not None:
                    att[bi, 0, (i + 1):, :(i + 1)] = 0
                if reset_position_ids:
                    position_ids[bi, (i + 1):] -= i + 1 - prev
                    prev = i + 1
    if att is not None:
        att = att < 0.5
    return att, loss_mask, position_ids


class MockGPTDataset(torch.utils.data.Dataset):
    def __init__(self, num_samples: int, seq_length: int, vocab_size: int, seed: int = 1234,
                 eod_token: int = 0, create_attention_mask: bool = False):
        self.num_samples, self.seq_length, self.vocab_size = num_samples, seq_length, vocab_size
        self.seed, self.eod, self.create_attention_mask = seed, eod_token, create_attention_mask

    def __len__(self):
        return self.num_samples

    def __getitem__(self, idx):
        rng = np.random.default_rng(self.seed + int(idx))
        toks = torch.from_numpy(rng.integers(1, self.vocab_size, size=self.seq_length + 1, dtype=np.int64))
        tokens, labels = toks[:-1].contiguous(), toks[1:].contiguous()
        out = {"tokens": tokens, "labels": labels, "loss_mask": torch.ones(self.seq_length, dtype=torch.float),
               "position_ids": torch.arange(self.seq_length, dtype=torch.long)}
        if self.create_attention_mask:
            out["attention_mask"] = torch.tril(torch.ones(self.seq_length, self.seq_length)).unsqueeze(0) < 0.5
        return out
