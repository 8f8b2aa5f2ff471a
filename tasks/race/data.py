"""RACE multiple-choice reading comprehension in BERT format (reference ``tasks/race/data.py``): every sample is
``[choices=4, seq]`` rows of ``[CLS] article [SEP] question (+) option [SEP]``; the article is trimmed to fit."""
from __future__ import annotations

import glob
import json
import os
import re
from typing import List

import torch

from tasks.glue.data import build_tokens_types_paddings

NUM_CHOICES = 4


class RaceDataset(torch.utils.data.Dataset):
    def __init__(self, dataset_name: str, datapaths: List[str], tokenizer, max_seq_length: int, max_qa_length: int = 128):
        self.name, self.tok, self.max_seq_length, self.max_qa_length = dataset_name, tokenizer, max_seq_length, max_qa_length
        self.samples = []
        for path in datapaths:
            files = sorted(glob.glob(os.path.join(path, "*.txt"))) if os.path.isdir(path) else [path]
            for fn in files:
                with open(fn) as f:
                    for line in f:
                        line = line.strip()
                        if line:
                            self._add(json.loads(line))

    def _add(self, data):
        context_ids = self.tok.tokenize(data["article"])
        for q, choices, ans in zip(data["questions"], data["options"], data["answers"]):
            label = ord(ans) - ord("A")
            rows = []
            for opt in choices[:NUM_CHOICES]:
                qa = re.sub(r"_+", " " + opt + " ", q) if "_" in q else q + " " + opt      # cloze vs plain question
                qa_ids = self.tok.tokenize(qa)[: self.max_qa_length]
                ids, types, pad = build_tokens_types_paddings(context_ids, qa_ids, self.max_seq_length, self.tok.cls,
                                                              self.tok.sep, self.tok.pad)
                rows.append((ids, types, pad))
            self.samples.append({"rows": rows, "label": label})

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        s = self.samples[i]
        return {"text": torch.tensor([r[0] for r in s["rows"]]), "types": torch.tensor([r[1] for r in s["rows"]]),
                "padding_mask": torch.tensor([r[2] for r in s["rows"]]), "label": torch.tensor(s["label"])}
