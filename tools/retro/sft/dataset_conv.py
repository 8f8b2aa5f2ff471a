"""Instruction / QA fine-tuning data for Retro (reference ``tools/retro/sft/dataset_conv.py``).

Input: a json list (or jsonl) of ``{"question": str, "answers": [str] | "answer": str, "ctxs": [{"title", "text"}, ...]}``
-- the retrieved passages come with the sample (DPR / open-instruct style files), they are *not* looked up at training
time.  A sample is ``User: {question}\\n\\nAssistant: {answer}<eod>`` padded to ``seq_length``; the loss covers the answer
only; every chunk of the sample sees the same ``num_neighbors`` passages (``title\\ntext`` truncated / padded to
``retrieved_length``) -- with ``ft_neighbours`` the first of them is also prepended to the prompt as plain context, the
reference's in-context variant."""
from __future__ import annotations

import json
from typing import List

import numpy as np
import torch


def read_samples(path: str) -> List[dict]:
    with open(path, encoding="utf-8") as f:
        head = f.read(1)
        f.seek(0)
        rows = json.load(f) if head == "[" else [json.loads(line) for line in f if line.strip()]
    out = []
    for r in rows:
        ans = r.get("answers") or r.get("answer") or r.get("output") or ""
        ans = ans[0] if isinstance(ans, (list, tuple)) and ans else ans
        out.append({"question": r.get("question") or r.get("instruction") or "", "answer": ans if isinstance(ans, str) else str(ans),
                    "ctxs": list(r.get("ctxs") or r.get("neighbours") or [])})
    return out


class RetroSFTDataset(torch.utils.data.Dataset):
    def __init__(self, path: str, tokenizer, seq_length: int, chunk_length: int, num_neighbors: int, retrieved_length: int,
                 ft_neighbours: int = 0, pad: int = 0, answer_loss_only: bool = True):
        self.samples = read_samples(path)
        self.tok, self.seq, self.chunk, self.k, self.r = tokenizer, seq_length, chunk_length, num_neighbors, retrieved_length
        self.ft_neighbours, self.pad, self.answer_loss_only = ft_neighbours, pad, answer_loss_only
        if seq_length % chunk_length:
            raise ValueError("seq_length must be a multiple of the retro chunk length")

    def __len__(self):
        return len(self.samples)

    def _ctx_tokens(self, ctx) -> List[int]:
        text = ctx if isinstance(ctx, str) else (f"{ctx.get('title', '')}\n{ctx.get('text', '')}" if ctx.get("title") else ctx.get("text", ""))
        return self.tok.tokenize(text)

    def __getitem__(self, i):
        s, tok = self.samples[i], self.tok
        prefix = ""
        for c in s["ctxs"][: self.ft_neighbours]:
            prefix += (c if isinstance(c, str) else c.get("text", "")) + "\n\n"
        prompt = tok.tokenize(f"{prefix}User: {s['question']}\n\nAssistant:")
        answer = tok.tokenize(" " + s["answer"]) + [tok.eod]
        budget = self.seq + 1
        if len(prompt) + len(answer) > budget:                    # keep the answer, trim the front of the prompt
            prompt = prompt[max(0, len(prompt) + len(answer) - budget):]
            answer = answer[: budget - len(prompt)]
        ids = prompt + answer
        text = np.full(budget, self.pad, dtype=np.int64)
        text[: len(ids)] = ids
        mask = np.zeros(budget, dtype=np.float32)
        lo = len(prompt) if self.answer_loss_only else 0
        mask[lo: len(ids)] = 1.0
        nb = np.full((self.seq // self.chunk, self.k, self.r), self.pad, dtype=np.int64)
        for j, c in enumerate(s["ctxs"][: self.k]):
            t = self._ctx_tokens(c)[: self.r]
            nb[:, j, : len(t)] = t
        return {"tokens": torch.from_numpy(text[:-1].copy()), "labels": torch.from_numpy(text[1:].copy()),
                "loss_mask": torch.from_numpy(mask[1:].copy()), "position_ids": torch.arange(self.seq),
                "neighbor_tokens": torch.from_numpy(nb)}
