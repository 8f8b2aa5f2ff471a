"""Supervised retriever fine-tuning on Natural Questions (reference ``tasks/orqa/supervised/{data,finetune,eval_utils}.py``,
task ``RET-FINETUNE-NQ``): DPR-style training of the bi-encoder -- every question must rank its positive passage first
among the positives *and hard negatives* of the whole data-parallel batch -- with an average-rank / top-k validation
on (positive + fixed negatives) lists after every epoch.

Data: the DPR json release (``[{"question", "answers", "positive_ctxs": [{"title", "text"}], "negative_ctxs": [...],
"hard_negative_ctxs": [...]}, ...]``)."""
from __future__ import annotations

import json
import math
import random
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def _pad(ids: List[int], n: int, cls: int, sep: int, pad: int):
    ids = [cls] + ids[: n - 2] + [sep]
    k = len(ids)
    return ids + [pad] * (n - k), [1] * k + [0] * (n - k)


def _ctx_ids(tok, ctx: Dict[str, str]) -> List[int]:
    return tok.tokenize(ctx.get("title", "")) + [tok.sep] + tok.tokenize(ctx["text"])


class NQSupervisedDataset(torch.utils.data.Dataset):
    def __init__(self, name: str, datapaths, tokenizer, max_seq_length: int, evaluate: bool = False, train_with_neg: bool = False,
                 train_hard_neg: int = 0, val_av_rank_hard_neg: int = 30, val_av_rank_other_neg: int = 30, sample_rate: float = 1.0,
                 seed: int = 1234):
        self.name, self.tok, self.n, self.evaluate = name, tokenizer, max_seq_length, evaluate
        self.train_with_neg, self.train_hard_neg = train_with_neg, train_hard_neg
        self.val_hard, self.val_other = val_av_rank_hard_neg, val_av_rank_other_neg
        self.rng = random.Random(seed)
        self.samples = []
        for path in ([datapaths] if isinstance(datapaths, str) else datapaths):
            with open(path, encoding="utf-8") as f:
                for row in json.load(f):
                    if not row.get("positive_ctxs"):
                        continue
                    q = row["question"]
                    self.samples.append({"question": q[:-1] if q.endswith("?") else q, "pos": row["positive_ctxs"][0],
                                         "hard": list(row.get("hard_negative_ctxs") or []), "neg": list(row.get("negative_ctxs") or []),
                                         "answers": row.get("answers", [])})
        if sample_rate < 1:
            self.samples = self.rng.sample(self.samples, int(len(self.samples) * sample_rate))

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        s, tok = self.samples[i], self.tok
        q, qm = _pad(tok.tokenize(s["question"]), self.n, tok.cls, tok.sep, tok.pad)
        c, cm = _pad(_ctx_ids(tok, s["pos"]), self.n, tok.cls, tok.sep, tok.pad)
        negs: List[dict] = []
        if self.evaluate:
            negs = s["neg"][: self.val_other] + s["hard"][: self.val_hard]
        elif self.train_with_neg:
            hard, other = list(s["hard"]), list(s["neg"])
            self.rng.shuffle(hard)
            self.rng.shuffle(other)
            negs = hard[: self.train_hard_neg]
            negs += other[: self.train_hard_neg - len(negs)]
        neg = [_pad(_ctx_ids(tok, c_), self.n, tok.cls, tok.sep, tok.pad) for c_ in negs]
        return {"query": q, "query_mask": qm, "context": c, "context_mask": cm, "neg": [x[0] for x in neg],
                "neg_mask": [x[1] for x in neg], "reference": s["answers"]}


def task_collate_fn(batch):
    t = lambda k: torch.tensor([b[k] for b in batch], dtype=torch.long)  # noqa: E731
    out = {k: t(k) for k in ("query", "query_mask", "context", "context_mask")}
    negs = [x for b in batch for x in b["neg"]]
    out["neg"] = torch.tensor(negs, dtype=torch.long) if negs else None
    out["neg_mask"] = torch.tensor([x for b in batch for x in b["neg_mask"]], dtype=torch.long) if negs else None
    out["reference"] = [b["reference"] for b in batch]
    return out


def _dp():
    from megatronapp_b200.parallel import state as ps

    if not (torch.distributed.is_initialized() and ps.model_parallel_is_initialized()):
        return None, 0, 1
    g = ps.get_data_parallel_group()
    if g is None or isinstance(g, ps._SingleGroup):
        return None, 0, 1
    return g, torch.distributed.get_rank(g), torch.distributed.get_world_size(g)


def _pad_rows_to_group_max(x: torch.Tensor, group) -> torch.Tensor:
    """ranks may hold different numbers of negatives: pad dim 0 to the group maximum so the embeddings all-gather"""
    n = torch.tensor([x.shape[0]], device=x.device)
    torch.distributed.all_reduce(n, op=torch.distributed.ReduceOp.MAX, group=group)
    return F.pad(x, (0, 0, 0, int(n) - x.shape[0])) if int(n) > x.shape[0] else x


def retrieval_forward_step(batch, model, score_scaling: bool = False, hidden_size: int = 1, train_with_neg: bool = False):
    from megatronapp_b200.legacy.model.biencoder_model import _AllGatherBatch

    dev = next(model.parameters()).device
    q, qm, c, cm = (batch[k].to(dev) for k in ("query", "query_mask", "context", "context_mask"))
    b = q.shape[0]
    group, rank, world = _dp()
    if batch["neg"] is not None:
        neg, nm = batch["neg"].to(dev), batch["neg_mask"].to(dev)
        if group is not None:
            neg, nm = _pad_rows_to_group_max(neg, group), _pad_rows_to_group_max(nm, group)
        c, cm = torch.cat([c, neg]), torch.cat([cm, nm])
    n_ctx = c.shape[0]
    ql, cl = model(q, qm, torch.zeros_like(q), c, cm, torch.zeros_like(c))

    def loss_fn(out):
        ql_, cl_ = out
        if group is not None and world > 1:
            ql_, cl_ = _AllGatherBatch.apply(ql_, group), _AllGatherBatch.apply(cl_, group)
        scores = ql_.float() @ cl_.float().t()
        if score_scaling:
            scores = scores / math.sqrt(hidden_size)
        # the positives of rank r sit at rows [r * n_ctx, r * n_ctx + b) of the gathered context matrix
        labels = torch.cat([torch.arange(r * n_ctx, r * n_ctx + b, device=dev) for r in range(world)])
        logp = F.log_softmax(scores, dim=1)
        loss = F.nll_loss(logp, labels)
        correct = (logp.argmax(1) == labels).float().sum()
        return loss, {"lm loss": loss.detach(), "correct_prediction_count": correct}

    return (ql, cl), loss_fn


@torch.no_grad()
def retrieval_metrics(model, loader, topk=(1, 5), score_scaling: bool = False, hidden_size: int = 1) -> Dict[str, float]:
    """Average rank of the positive passage and top-k accuracy over ``[positive of every question in the batch] + [the
    batch's fixed negatives]`` (reference ``eval_utils.retrieval_loss``)."""
    was_training = model.training
    model.eval()
    dev = next(model.parameters()).device
    stats = {"rank": 0.0, **{f"top{k}_acc": 0.0 for k in topk}}
    total = 0
    for batch in loader:
        q, qm, c, cm = (batch[k].to(dev) for k in ("query", "query_mask", "context", "context_mask"))
        if batch["neg"] is not None:
            c, cm = torch.cat([c, batch["neg"].to(dev)]), torch.cat([cm, batch["neg_mask"].to(dev)])
        ql, cl = model(q, qm, torch.zeros_like(q), c, cm, torch.zeros_like(c))
        scores = ql.float() @ cl.float().t()
        if score_scaling:
            scores = scores / math.sqrt(hidden_size)
        labels = torch.arange(q.shape[0], device=dev)
        rank = (scores > scores.gather(1, labels[:, None])).sum(1)
        stats["rank"] += float(rank.sum())
        for k in topk:
            stats[f"top{k}_acc"] += 100.0 * float((rank < k).sum())
        total += q.shape[0]
    t = torch.tensor([total] + [stats[k] for k in sorted(stats)], dtype=torch.float64, device=dev)
    group, _, world = _dp()
    if group is not None and world > 1:
        torch.distributed.all_reduce(t, group=group)
    model.train(was_training)
    out = {k: float(v) / max(float(t[0]), 1.0) for k, v in zip(sorted(stats), t[1:])}
    out["total"] = int(t[0])
    return out


def main():
    from functools import partial

    from megatronapp_b200.training.global_vars import get_args, get_tokenizer
    from megatronapp_b200.training.utils import print_rank_0
    from pretrain_ict import model_provider
    from tasks.finetune_utils import finetune

    args = get_args()
    tok = get_tokenizer()
    model = model_provider()
    if torch.cuda.is_available():
        model = model.cuda()
    if args.load:
        from megatronapp_b200.training import checkpointing

        checkpointing.load_checkpoint([model], None, None)
    kw = dict(train_with_neg=args.train_with_neg, train_hard_neg=args.train_hard_neg, val_av_rank_hard_neg=args.val_av_rank_hard_neg,
              val_av_rank_other_neg=args.val_av_rank_other_neg)
    train = NQSupervisedDataset("training", args.train_data, tok, args.retriever_seq_length, evaluate=False,
                                sample_rate=args.sample_rate, **kw)
    valid = NQSupervisedDataset("validation", args.valid_data, tok, args.retriever_seq_length, evaluate=True, **kw)
    topk = tuple(args.retriever_report_topk_accuracies or (1, 5))
    fwd = partial(retrieval_forward_step, score_scaling=args.retriever_score_scaling, hidden_size=args.hidden_size,
                  train_with_neg=args.train_with_neg)
    ev = partial(retrieval_metrics, topk=topk, score_scaling=args.retriever_score_scaling, hidden_size=args.hidden_size)
    hist = finetune(model, train, valid, epochs=args.epochs, batch_size=args.micro_batch_size, lr=args.lr, forward_step=fwd,
                    evaluate=ev, collate_fn=task_collate_fn, eval_batch_size=args.eval_micro_batch_size or args.micro_batch_size)
    for h in hist:
        print_rank_0("epoch:{}".format(h["epoch"]) + "".join(f"|{k} = {v:.2f}" for k, v in h.items() if k not in ("epoch", "total")))
    if args.save:
        from megatronapp_b200.training import checkpointing

        checkpointing.save_checkpoint(args.epochs, [model], None, None)
    return hist
