"""Wikipedia evidence for open-retrieval QA (reference ``legacy/data/orqa_wiki_dataset.py``): the DPR TSV
(``id \\t text \\t title``), one passage per row, tokenised as ``[CLS] title [SEP] text [SEP]`` and padded."""
from __future__ import annotations

import csv
from typing import Dict, List

import numpy as np
import torch


class OpenRetrievalEvidenceDataset(torch.utils.data.Dataset):
    def __init__(self, task_name: str, dataset_name: str, datapath: str, tokenizer, max_seq_length: int, sample_rate: float = 1.0):
        self.task_name, self.dataset_name, self.tok, self.n = task_name, dataset_name, tokenizer, max_seq_length
        self.samples, self.id2text = self.process_samples_from_single_path(datapath)
        if sample_rate < 1:
            rng = np.random.RandomState(1234)
            keep = sorted(rng.choice(len(self.samples), int(len(self.samples) * sample_rate), replace=False).tolist())
            self.samples = [self.samples[i] for i in keep]

    @staticmethod
    def process_samples_from_single_path(filename: str):
        samples, id2text = [], {}
        with open(filename, newline="", encoding="utf-8") as f:
            for i, row in enumerate(csv.reader(f, delimiter="\t")):
                if i == 0 and row and row[0] == "id":
                    continue
                doc_id, text, title = int(row[0]), row[1], (row[2] if len(row) > 2 else "")
                samples.append({"doc_id": doc_id, "text": text, "title": title})
                id2text[doc_id] = (text, title)
        return samples, id2text

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i) -> Dict[str, np.ndarray]:
        s, tok = self.samples[i], self.tok
        ids = [tok.cls] + tok.tokenize(s["title"]) + [tok.sep] + tok.tokenize(s["text"])
        ids = ids[: self.n - 1] + [tok.sep]
        k = len(ids)
        return {"row_id": np.int64(s["doc_id"]), "context": np.asarray(ids + [tok.pad] * (self.n - k), dtype=np.int64),
                "context_mask": np.asarray([1] * k + [0] * (self.n - k), dtype=np.int64),
                "context_types": np.zeros(self.n, dtype=np.int64), "context_pad_mask": np.asarray([1] * k + [0] * (self.n - k), dtype=np.int64)}


def get_open_retrieval_wiki_dataset(evidence_data_path: str, tokenizer, retriever_seq_length: int, sample_rate: float = 1.0):
    return OpenRetrievalEvidenceDataset("ICT-ZEROSHOT-NQ", "evidence", evidence_data_path, tokenizer, retriever_seq_length, sample_rate)


def get_open_retrieval_batch(batch: List[Dict[str, np.ndarray]]):
    t = lambda k: torch.from_numpy(np.stack([b[k] for b in batch]))  # noqa: E731
    return t("row_id"), t("context"), t("context_mask"), t("context_types"), t("context_pad_mask")
