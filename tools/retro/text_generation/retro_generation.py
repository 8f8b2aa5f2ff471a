"""Retrieval-augmented generation with a Retro model (reference ``tools/retro/text_generation/{retro_generation,retro_api,
retro_text_generation}.py``).

The prompt is cut into ``chunk_length`` windows; every window queries the project's index (``Retriever``) for its
neighbours -- chunk + continuation, ``retrieved_length`` tokens each -- and the decoder attends to them through chunked
cross-attention.  While generating, whenever the sequence grows into a new chunk, that chunk's neighbours are retrieved
from what has been generated so far (the reference's ``--retro-gen-*`` behaviour: retrieval keeps up with the text).
Decoding is greedy or top-k / top-p; every step re-runs the model on the whole sequence (no KV cache: the neighbour set of
the last chunk changes as it fills up).

    python tools/retro/text_generation/retro_generation.py --retro-project-dir work --load ckpt --prompts questions.txt ...
"""
from __future__ import annotations

import os
import sys
from typing import Callable, List, Optional

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))


class Retriever:
    """Neighbour lookup against a preprocessed Retro project (chunk DB + index + the embedder used to build it)."""

    def __init__(self, project_dir: str, num_neighbors: Optional[int] = None, nprobe: Optional[int] = None):
        from megatronapp_b200.data.retro import ChunkDB, RetroPreprocessingConfig
        from megatronapp_b200.data.retro.index import load_index
        from megatronapp_b200.data.retro.query import get_embedder

        self.cfg = RetroPreprocessingConfig.load(project_dir)
        self.db = ChunkDB(self.cfg)
        self.index = load_index(self.cfg.path("index", "index.npz"))
        self.embed = get_embedder(self.cfg)
        self.k = num_neighbors or self.cfg.retro_query_num_neighbors_save
        self.nprobe = nprobe or self.cfg.retro_query_nprobe

    def __call__(self, chunks: np.ndarray) -> np.ndarray:
        """``chunks [n, chunk_length]`` token ids (pad with 0) -> ``[n, k, retrieved_length]`` neighbour tokens."""
        _, ids = self.index.search(self.embed(chunks), self.k, self.nprobe)
        R = self.cfg.retro_retrieved_length
        out = np.zeros((chunks.shape[0], self.k, R), dtype=np.int64)
        for i in range(chunks.shape[0]):
            for j, cid in enumerate(ids[i][: self.k]):
                if cid >= 0:
                    out[i, j] = self.db.tokens(int(cid), R)
        return out


def _sample(logits: torch.Tensor, top_k: int, top_p: float, temperature: float, generator=None) -> int:
    if top_k == 1 or temperature == 0.0:
        return int(logits.argmax())
    logits = logits.float() / max(temperature, 1e-6)
    if top_k > 0:
        kth = logits.topk(min(top_k, logits.numel())).values[-1]
        logits = logits.masked_fill(logits < kth, float("-inf"))
    probs = torch.softmax(logits, -1)
    if 0.0 < top_p < 1.0:
        sp, si = probs.sort(descending=True)
        keep = sp.cumsum(0) - sp <= top_p
        probs = torch.zeros_like(probs).scatter_(0, si[keep], sp[keep])
        probs = probs / probs.sum()
    return int(torch.multinomial(probs, 1, generator=generator))


@torch.no_grad()
def retro_generate(model, prompt_tokens: List[int], retrieve: Callable[[np.ndarray], np.ndarray], chunk_length: int,
                   max_new_tokens: int = 32, eod: Optional[int] = None, top_k: int = 1, top_p: float = 0.0,
                   temperature: float = 1.0, pad: int = 0, max_seq_length: Optional[int] = None, generator=None) -> List[int]:
    """Returns the generated continuation (without the prompt)."""
    model.eval()
    dev = next(model.parameters()).device
    toks = list(prompt_tokens)
    cache = {}                                           # chunk index -> neighbours of a COMPLETE chunk (never changes)
    out: List[int] = []
    for _ in range(max_new_tokens):
        n = len(toks)
        n_chunks = -(-n // chunk_length)
        padded = np.full(n_chunks * chunk_length, pad, dtype=np.int64)
        padded[:n] = toks
        chunks = padded.reshape(n_chunks, chunk_length)
        need = [c for c in range(n_chunks) if c not in cache]
        got = retrieve(chunks[need]) if need else None
        nb = []
        for c in range(n_chunks):
            if c in cache:
                nb.append(cache[c])
                continue
            x = got[need.index(c)]
            if (c + 1) * chunk_length <= n:
                cache[c] = x                             # the chunk is full: its neighbours are final
            nb.append(x)
        ids = torch.from_numpy(padded).to(dev).unsqueeze(0)
        pos = torch.arange(ids.shape[1], device=dev).unsqueeze(0)
        ctx = torch.from_numpy(np.stack(nb)).to(dev).unsqueeze(0)               # [1, chunks, k, r]
        logits = model(ids, pos, None, context_input_ids=ctx)
        if logits.dim() == 3 and logits.shape[0] != 1:                           # [s, b, v] -> [b, s, v]
            logits = logits.transpose(0, 1)
        new = _sample(logits[0, n - 1], top_k, top_p, temperature, generator)
        toks.append(new)
        out.append(new)
        if (eod is not None and new == eod) or (max_seq_length and len(toks) >= max_seq_length):
            break
    return out


def main():
    from megatronapp_b200.training import checkpointing
    from megatronapp_b200.training.global_vars import get_args, get_tokenizer
    from megatronapp_b200.training.initialize import initialize_megatron
    from megatronapp_b200.training.training import get_model
    from pretrain_retro import add_retro_args, model_provider

    def extra(parser):
        parser = add_retro_args(parser)
        g = parser.add_argument_group("retro generation")
        g.add_argument("--prompts", required=True, help="text file, one prompt per line")
        g.add_argument("--output", default="retro_generations.txt")
        g.add_argument("--out-seq-length", type=int, default=64)
        g.add_argument("--top-k", type=int, default=1)
        g.add_argument("--top-p", type=float, default=0.0)
        g.add_argument("--temperature", type=float, default=1.0)
        return parser

    initialize_megatron(extra_args_provider=extra, args_defaults={"tokenizer_type": "GPT2BPETokenizer"})
    args = get_args()
    tok = get_tokenizer()
    model = get_model(model_provider, wrap_with_ddp=False)
    if args.load:
        checkpointing.load_checkpoint(model, None, None)
    retrieve = Retriever(args.retro_project_dir, num_neighbors=args.retro_num_neighbors)
    with open(args.prompts, encoding="utf-8") as f, open(args.output, "w", encoding="utf-8") as g:
        for line in f:
            prompt = line.rstrip("\n")
            gen = retro_generate(model[0], tok.tokenize(prompt), retrieve, args.retro_chunk_length, args.out_seq_length,
                                 eod=tok.eod, top_k=args.top_k, top_p=args.top_p, temperature=args.temperature,
                                 max_seq_length=args.seq_length)
            text = tok.detokenize([t for t in gen if t != tok.eod])
            g.write(text.replace("\n", " ") + "\n")
            print(f"{prompt} -> {text}", flush=True)


if __name__ == "__main__":
    main()
