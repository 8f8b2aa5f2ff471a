"""Zero-shot GPT evaluation: WikiText-103 perplexity (sliding window with ``--overlapping-eval``
stride, token -> word perplexity adjustment) and LAMBADA last-word accuracy.

Parity: reference ``tasks/zeroshot_gpt/{evaluate,datasets,detokenizer}.py``.
"""
from __future__ import annotations

import math
import re
from typing import List

import torch


# ------------------------------------------------------------------------------ datasets
def wikitext_detokenizer(s: str) -> str:
    s = s.replace("s '", "s'")
    s = re.sub(r"/' [0-9]/", r"/'[0-9]/", s)
    for a, b in ((" @-@ ", "-"), (" @,@ ", ","), (" @.@ ", "."), (" : ", ": "), (" ; ", "; "), (" . ", ". "),
                 (" ! ", "! "), (" ? ", "? "), (" , ", ", ")):
        s = s.replace(a, b)
    s = re.sub(r"\(\s*([^\)]*?)\s*\)", r"(\1)", s)
    s = re.sub(r"\[\s*([^\]]*?)\s*\]", r"[\1]", s)
    s = re.sub(r"{\s*([^}]*?)\s*}", r"{\1}", s)
    s = re.sub(r"\"\s*([^\"]*?)\s*\"", r'"\1"', s)
    s = re.sub(r"'\s*([^']*?)\s*'", r"'\1'", s)
    for a, b in (("= = = =", "===="), ("= = =", "==="), ("= =", "=="), (" " + chr(176) + " ", chr(176)),
                 (" \n", "\n"), ("\n ", "\n"), (" N ", " 1 "), (" 's", "'s")):
        s = s.replace(a, b)
    return s


class LMDataset(torch.utils.data.Dataset):
    """Sliding windows of ``seq_len`` tokens; only the last ``overlapping_eval`` targets of every window
    after the first are scored, so each token is predicted once with maximal context."""

    def __init__(self, tokens: List[int], seq_len: int, pad_idx: int, num_original_tokens: int,
                 num_tokenized_tokens: int, overlapping_eval: int = None):
        self.tokens, self.seq_len, self.pad_idx = tokens, seq_len, pad_idx
        self.overlapping_eval = max(1, overlapping_eval or seq_len)
        self.num_original_tokens, self.num_tokenized_tokens = num_original_tokens, num_tokenized_tokens
        targets = max(len(tokens) - 1 - self.overlapping_eval, 0)
        self.total_sequences = max(math.ceil(targets / self.overlapping_eval) + 1, 1)

    def __len__(self):
        return self.total_sequences

    def __getitem__(self, idx):
        start = idx * self.overlapping_eval
        end = start + self.seq_len
        toks = self.tokens[start:end + 1]
        n = len(toks)
        pad_mask = [1] * n
        if n < self.seq_len + 1:
            pad_mask += [0] * (self.seq_len + 1 - n)
            toks = toks + [self.pad_idx] * (self.seq_len + 1 - n)
        pad_mask = torch.tensor(pad_mask[1:], dtype=torch.float32)
        if self.overlapping_eval != self.seq_len and idx != 0:
            pad_mask[:-self.overlapping_eval] = 0
        return {"text": torch.tensor(toks, dtype=torch.int64), "pad_mask": pad_mask}


class LambadaDataset(torch.utils.data.Dataset):
    def __init__(self, path: str, pad_idx: int, tokenizer, seq_len: int, strict: bool = False):
        import json

        self.seq_len, self.pad_idx, self.tok, self.strict = seq_len, pad_idx, tokenizer, strict
        self.tokens, self.labels = [], []
        with open(path) as f:
            for line in f:
                text = json.loads(line)["text"]
                t, lab = self.get_tokens(text)
                self.tokens.append(t)
                self.labels.append(lab)

    def get_tokens(self, text):
        if not self.strict:
            toks = self.tok.tokenize(text)
            return toks[:-1], [toks[-1]]
        last = text.split()[-1]
        start = text.rfind(last)
        return self.tok.tokenize(text[:start].strip()), self.tok.tokenize(" " + last)

    def __len__(self):
        return len(self.tokens)

    def __getitem__(self, idx):
        toks, lab = self.tokens[idx], self.labels[idx]
        pad_mask = [0] * len(toks) + [1] * len(lab)
        toks = toks + lab
        n = len(toks)
        if n < self.seq_len + 1:
            pad_mask += [0] * (self.seq_len + 1 - n)
            toks = toks + [self.pad_idx] * (self.seq_len + 1 - n)
        return {"text": torch.tensor(toks[:self.seq_len + 1], dtype=torch.int64),
                "pad_mask": torch.tensor(pad_mask[1:self.seq_len + 1], dtype=torch.float32)}


def build_dataset(task: str, path: str, tokenizer, seq_len: int, overlapping_eval=None, strict_lambada=False):
    if task == "LAMBADA":
        return LambadaDataset(path, tokenizer.eod, tokenizer, seq_len, strict_lambada)
    with open(path, "rb") as f:
        raw = f.read().decode("utf-8")
    n_orig = len(raw.strip().split(" "))
    toks = tokenizer.tokenize(wikitext_detokenizer(raw))
    return LMDataset(toks, seq_len, tokenizer.eod, n_orig, len(toks), overlapping_eval)


# ------------------------------------------------------------------------------ evaluation
def _pipeline_forward(model, inp, pos, labels):
    """One forward through all pipeline stages; returns the model output on the last stage and ``None`` elsewhere."""
    from megatronapp_b200.parallel import p2p, state as ps

    if not ps.model_parallel_is_initialized() or ps.get_pipeline_model_parallel_world_size() == 1:
        return model(inp, pos, None, labels=labels)
    inner = model
    while hasattr(inner, "module"):
        inner = inner.module
    cfg = inner.config
    if not ps.is_pipeline_first_stage():
        h = p2p.recv((inp.shape[1], inp.shape[0], cfg.hidden_size), cfg.pipeline_dtype or cfg.params_dtype,
                     ps.get_pipeline_model_parallel_prev_rank(), "fwd", "recv-forward")
        inner.set_input_tensor(h)
    out = model(inp, pos, None, labels=labels if ps.is_pipeline_last_stage() else None)
    if not ps.is_pipeline_last_stage():
        p2p.send(out, ps.get_pipeline_model_parallel_next_rank(), "fwd", "send-forward")
        p2p.flush()
        return None
    return out


@torch.no_grad()
def evaluate(model, loader, eval_metric: str, device=None) -> float:
    """``loss``: summed token NLL over scored positions; ``accuracy``: number of samples whose every
    scored position is predicted correctly (LAMBADA).  Batches are dealt round-robin to the data-parallel replicas,
    pipeline stages pass activations along, and every rank returns the global total."""
    import torch.distributed as dist

    from megatronapp_b200.parallel import state as ps
    from megatronapp_b200.parallel.mappings import gather_from_tensor_model_parallel_region

    model.eval()
    total = 0.0
    dev = device or next(model.parameters()).device
    mp = ps.model_parallel_is_initialized()
    dp_rank, dp_size = (ps.get_data_parallel_rank(), ps.get_data_parallel_world_size()) if mp else (0, 1)
    for i, batch in enumerate(loader):
        if i % dp_size != dp_rank:
            continue
        toks = batch["text"].to(dev)
        mask = batch["pad_mask"].to(dev)
        inp, labels = toks[:, :-1].contiguous(), toks[:, 1:].contiguous()
        pos = torch.arange(inp.shape[1], device=dev).unsqueeze(0).expand_as(inp)
        if eval_metric == "loss":
            loss = _pipeline_forward(model, inp, pos, labels)
            if loss is not None:
                total += float((loss.float() * mask).sum())
        else:
            logits = _pipeline_forward(model, inp, pos, None)
            if logits is None:
                continue
            logits = logits.float()
            if mp and ps.get_tensor_model_parallel_world_size() > 1:
                logits = gather_from_tensor_model_parallel_region(logits)
            correct = (logits.argmax(-1) == labels).float()
            correct[mask == 0] = 1.0
            total += float(correct.prod(-1).sum())
    if mp and dist.is_initialized():
        t = torch.tensor([total], device=dev, dtype=torch.float64)
        for g in (ps.get_data_parallel_group(), ps.get_pipeline_model_parallel_group()):
            # DP: partial sums of the replicas; PP: only the last stage counted anything
            if g is not None and not isinstance(g, ps._SingleGroup):
                dist.all_reduce(t, group=g)
        total = float(t)
    return total


def summarize(task: str, output: float, dataset) -> dict:
    if task == "WIKITEXT103":
        n_tok, n_orig = dataset.num_tokenized_tokens, dataset.num_original_tokens
        val_loss = output / max(n_tok - 1, 1)
        ppl = math.exp(min(20, val_loss))
        ratio = (n_tok - 1) / max(n_orig - 1, 1)
        return {"avg loss": val_loss, "ppl": ppl, "adjusted ppl": math.exp(min(20, val_loss * ratio)), "token ratio": ratio}
    n = len(dataset)
    return {"number correct": output, "total examples": n, "avg accuracy": output / max(n, 1)}


def main():
    from megatronapp_b200.data.samplers import build_pretraining_data_loader  # noqa: F401
    from megatronapp_b200.training import checkpointing
    from megatronapp_b200.training.global_vars import get_args, get_tokenizer
    from megatronapp_b200.training.training import get_model
    from megatronapp_b200.training.utils import print_rank_0
    from pretrain_gpt import model_provider

    args = get_args()
    model = get_model(model_provider, wrap_with_ddp=False)
    if args.load:
        checkpointing.load_checkpoint(model, None, None)
    tok = get_tokenizer()
    ds = build_dataset(args.task, args.valid_data[0], tok, args.seq_length, args.overlapping_eval, args.strict_lambada)
    bs = args.eval_micro_batch_size or args.micro_batch_size
    loader = torch.utils.data.DataLoader(ds, batch_size=bs, shuffle=False, drop_last=False)
    out = evaluate(model[0], loader, "loss" if args.task == "WIKITEXT103" else "accuracy")
    res = summarize(args.task, out, ds)
    print_rank_0(" | ".join(f"{k}: {v:.4E}" if isinstance(v, float) else f"{k}: {v}" for k, v in res.items()))
    return res
