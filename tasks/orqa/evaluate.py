"""Unsupervised open-retrieval QA evaluation (reference ``tasks/orqa/{evaluate_orqa,evaluate_utils}.py``,
``ORQAEvaluator``): embed the evidence passages with the context tower, embed the questions with the query tower,
take the top-k passages by inner product, report the fraction of questions whose top-k contains an answer.

The reference shards a FAISS index over ranks; here the index is exact: the evidence matrix is split over the
data-parallel ranks, every rank scores all questions against its shard on the device (one GEMM per batch) and the
per-shard top-k lists are merged."""
import csv
from typing import Dict, List, Sequence, Tuple

import torch

from .qa_utils import calculate_matches


def read_nq_questions(path: str) -> Tuple[List[str], List[List[str]]]:
    """TSV: ``question \\t ["answer 1", "answer 2", ...]`` (the NQ-open files used by DPR / the reference)."""
    import ast

    qs, ans = [], []
    with open(path, newline="") as f:
        for row in csv.reader(f, delimiter="\t"):
            if len(row) >= 2:
                qs.append(row[0])
                ans.append(list(ast.literal_eval(row[1])))
    return qs, ans


def read_evidence(path: str) -> Dict[int, Tuple[str, str]]:
    """TSV with header ``id \\t text \\t title`` (DPR wikipedia split format)."""
    docs = {}
    with open(path, newline="") as f:
        for i, row in enumerate(csv.reader(f, delimiter="\t")):
            if i == 0 and row and row[0] == "id":
                continue
            docs[int(row[0])] = (row[1], row[2] if len(row) > 2 else "")
    return docs


class ORQAEvaluator:
    def __init__(self, model, tokenizer, seq_length: int = 64, batch_size: int = 128, device=None):
        self.model, self.tok, self.seq_length, self.batch_size = model, tokenizer, seq_length, batch_size
        self.device = device or next(model.parameters()).device
        self.doc_ids: List[int] = []
        self.doc_emb: torch.Tensor = torch.empty(0)

    def _encode(self, texts: Sequence[Tuple[str, str]], query: bool) -> torch.Tensor:
        tok, out = self.tok, []
        self.model.eval()
        for i in range(0, len(texts), self.batch_size):
            rows, masks = [], []
            for text, title in texts[i:i + self.batch_size]:
                ids = [tok.cls] + (tok.tokenize(title) + [tok.sep] if title else []) + tok.tokenize(text)
                ids = ids[: self.seq_length - 1] + [tok.sep]
                rows.append(ids + [tok.pad] * (self.seq_length - len(ids)))
                masks.append([1] * len(ids) + [0] * (self.seq_length - len(ids)))
            t, m = torch.tensor(rows, device=self.device), torch.tensor(masks, device=self.device)
            with torch.no_grad():
                out.append(self.model.embed_text(t, m, torch.zeros_like(t), query=query).float())
        return torch.cat(out) if out else torch.empty(0)

    def index_evidence(self, docs: Dict[int, Tuple[str, str]]):
        """This rank indexes ``docs[rank::world]``."""
        dist = torch.distributed
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
        self.doc_ids = sorted(docs)[rank::world]
        self.doc_emb = self._encode([docs[d] for d in self.doc_ids], query=False)

    def search(self, questions: List[str], top_k: int) -> List[Tuple[List[int], List[float]]]:
        q = self._encode([(x, "") for x in questions], query=True)
        k = min(top_k, len(self.doc_ids))
        scores, idx = (q @ self.doc_emb.t()).topk(k, dim=1)
        ids = torch.tensor(self.doc_ids, device=idx.device)[idx]
        dist = torch.distributed
        if dist.is_initialized() and dist.get_world_size() > 1:
            pairs = [None] * dist.get_world_size()
            dist.all_gather_object(pairs, (scores.cpu(), ids.cpu()))
            scores = torch.cat([p[0] for p in pairs], 1)
            ids = torch.cat([p[1] for p in pairs], 1)
            scores, order = scores.topk(min(top_k, scores.shape[1]), dim=1)
            ids = ids.gather(1, order)
        return [(ids[i].tolist(), scores[i].tolist()) for i in range(len(questions))]

    def evaluate(self, questions: List[str], answers: List[List[str]], docs: Dict[int, Tuple[str, str]],
                 report_topk: Sequence[int] = (1, 5, 20, 100), match: str = "string") -> Dict[int, float]:
        closest = self.search(questions, max(report_topk))
        hits, _ = calculate_matches(docs, answers, closest, match)
        out = {k: 100.0 * hits[min(k, len(hits)) - 1] / len(questions) for k in report_topk if hits}
        for k, v in out.items():
            print(f"top-{k} retrieval accuracy: {v:.2f}")
        return out
