"""GLUE sentence-pair datasets (MNLI, QQP) in BERT input format (reference ``tasks/glue/{data,mnli,qqp}.py`` and
``tasks/data_utils.py build_tokens_types_paddings_from_text``)."""
from __future__ import annotations

import csv
from typing import Dict, List

import torch

LABELS = {"MNLI": {"contradiction": 0, "entailment": 1, "neutral": 2}, "QQP": {"0": 0, "1": 1}}


def build_tokens_types_paddings(text_a_ids: List[int], text_b_ids: List[int], max_seq_length: int, cls_id: int, sep_id: int,
                                pad_id: int):
    a, b = list(text_a_ids), list(text_b_ids or [])
    budget = max_seq_length - (3 if b else 2)
    while len(a) + len(b) > budget:                     # trim the longer segment from its tail
        (a if len(a) > len(b) else b).pop()
    ids = [cls_id] + a + [sep_id] + (b + [sep_id] if b else [])
    types = [0] * (len(a) + 2) + [1] * (len(b) + 1 if b else 0)
    n = len(ids)
    pad = max_seq_length - n
    return ids + [pad_id] * pad, types + [pad_id] * pad, [1] * n + [0] * pad


class GLUEDataset(torch.utils.data.Dataset):
    def __init__(self, task: str, name: str, datapaths: List[str], tokenizer, max_seq_length: int):
        self.task, self.name, self.tok, self.max_seq_length = task, name, tokenizer, max_seq_length
        self.samples: List[Dict] = []
        for p in datapaths:
            self.samples.extend(self._read(p))

    def _read(self, path):
        out = []
        with open(path, newline="") as f:
            rows = list(csv.reader(f, delimiter="\t", quoting=csv.QUOTE_NONE))
        header, rows = rows[0], rows[1:]
        col = {h: i for i, h in enumerate(header)}
        for r in rows:
            if self.task == "MNLI":
                a, b, lab, uid = r[col["sentence1"]], r[col["sentence2"]], r[col.get("gold_label", len(r) - 1)], r[col.get("index", 0)]
            else:
                a, b, lab, uid = r[col["question1"]], r[col["question2"]], r[col.get("is_duplicate", len(r) - 1)], r[col.get("id", 0)]
            if lab not in LABELS[self.task]:
                continue
            out.append({"text_a": a.strip(), "text_b": b.strip(), "label": LABELS[self.task][lab], "uid": int(uid) if str(uid).isdigit() else len(out)})
        return out

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        s = self.samples[i]
        ids, types, pad = build_tokens_types_paddings(self.tok.tokenize(s["text_a"]), self.tok.tokenize(s["text_b"]),
                                                      self.max_seq_length, self.tok.cls, self.tok.sep, self.tok.pad)
        return {"text": torch.tensor(ids), "types": torch.tensor(types), "padding_mask": torch.tensor(pad),
                "label": torch.tensor(s["label"]), "uid": torch.tensor(s["uid"])}
