"""Helper functions of the reference's ``megatron/inference/text_generation/{sampling,tokenization,communication}.py``."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from ..parallel import state as ps


# -- sampling.py -----------------------------------------------------------------------------------------------------------
def modify_logits_for_top_k_filtering(logits: torch.Tensor, top_k: int) -> None:
    """In place: everything below the k-th largest logit of a row becomes -inf."""
    kth = torch.topk(logits, top_k)[0][..., -1, None]
    logits.masked_fill_(logits < kth, float("-Inf"))


def modify_logits_for_top_p_filtering(logits: torch.Tensor, top_p: float) -> None:
    """In place nucleus filter: the smallest set of tokens whose probability mass reaches ``top_p`` survives."""
    sorted_logits, sorted_idx = torch.sort(logits, descending=True)
    cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
    drop = cum > top_p
    drop[..., 1:] = drop[..., :-1].clone()          # keep the first token that crosses the threshold
    drop[..., 0] = False
    logits.masked_fill_(drop.scatter(1, sorted_idx, drop), float("-Inf"))


def sample(logits: torch.Tensor, top_k: int = 0, top_p: float = 0.0, temperature: float = 1.0, vocab_size: Optional[int] = None):
    """``logits [b, v]`` -> ``[b]`` sampled ids (greedy for ``top_k == 1``); ids are clamped below ``vocab_size`` (padding)."""
    if logits.dim() != 2:
        raise ValueError("expected the logits to be of [b, v] size")
    if top_k == 1:
        if top_p != 0.0:
            raise ValueError("cannot set both greedy and top-p samplings")
        out = torch.argmax(logits, dim=-1)
    else:
        logits = logits.float().clone()
        if temperature != 1.0:
            logits.div_(temperature)
        if top_k > 1:
            if top_p != 0.0:
                raise ValueError("cannot set both top-k and top-p samplings")
            if top_k > logits.size(1):
                raise ValueError("top-k is larger than logit size")
            modify_logits_for_top_k_filtering(logits, top_k)
        elif top_p > 0.0:
            if top_p > 1.0:
                raise ValueError("top-p should be in (0, 1]")
            modify_logits_for_top_p_filtering(logits, top_p)
        out = torch.multinomial(logits.softmax(dim=-1), num_samples=1).view(-1)
    if vocab_size:
        out = torch.clamp(out, min=0, max=vocab_size - 1)
    return out


# -- tokenization.py ---------------------------------------------------------------------------------------------------------
def _tok():
    from ..training.global_vars import get_tokenizer

    return get_tokenizer()


def detokenize_generations(tokens_gpu_tensor, lengths_gpu_tensor, detokenize_segments: bool, tokenizer=None):
    """-> (tokens as lists, generations as strings[, per-token segments])."""
    tok = tokenizer or _tok()
    tokens = tokens_gpu_tensor.cpu().tolist()
    lengths = lengths_gpu_tensor.cpu().tolist()
    texts, segments = [], []
    for row, n in zip(tokens, lengths):
        row = row[:n]
        texts.append(tok.detokenize(row))
        if detokenize_segments:
            segments.append([tok.detokenize([t]) for t in row])
    return (tokens, texts, segments) if detokenize_segments else (tokens, texts)


def tokenize_prompts(prompts: Optional[List[str]] = None, tokens_to_generate: Optional[int] = None, add_BOS: Optional[bool] = None,
                     rank: int = 0, data_parallel: bool = False, tokenizer=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Tokenise on ``rank`` and broadcast: ``(tokens [b, max_prompt + tokens_to_generate] padded with eod, lengths [b])``."""
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    me = dist.get_rank() if dist.is_initialized() else 0
    tokens = lengths = None
    sizes = [0, 0]
    if me == rank:
        tok = tokenizer or _tok()
        rows = [([tok.bos] if add_BOS and getattr(tok, "bos", None) is not None else []) + list(tok.tokenize(p)) for p in prompts]
        lens = [len(r) for r in rows]
        total = max(lens) + int(tokens_to_generate or 0)
        tokens = torch.tensor([r + [tok.eod] * (total - len(r)) for r in rows], dtype=torch.long, device=dev)
        lengths = torch.tensor(lens, dtype=torch.long, device=dev)
        sizes = [tokens.size(0), tokens.size(1)]
    sizes = broadcast_int_list(2, int_list=sizes, rank=rank)
    tokens = broadcast_tensor((sizes[0], sizes[1]), torch.long, tensor=tokens, rank=rank)
    lengths = broadcast_tensor((sizes[0],), torch.long, tensor=lengths, rank=rank)
    return tokens, lengths


# -- communication.py ------------------------------------------------------------------------------------------------------
def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def _real(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def broadcast_tensor(size, dtype, tensor=None, rank: int = 0, data_parallel: bool = False):
    """Every rank gets ``tensor`` of ``rank`` (shape / dtype must be known everywhere)."""
    me = dist.get_rank() if dist.is_initialized() else 0
    if me == rank:
        if tensor is None or tuple(tensor.shape) != tuple(size) or tensor.dtype != dtype:
            raise ValueError("the source rank must pass a tensor of the announced size and dtype")
        buf = tensor.to(_dev()).contiguous()
    else:
        buf = torch.empty(size, dtype=dtype, device=_dev())
    if _real():
        dist.broadcast(buf, rank)
    return buf


def broadcast_list(size: int, dtype, list_values=None, rank: int = 0, data_parallel: bool = False):
    me = dist.get_rank() if dist.is_initialized() else 0
    t = torch.tensor(list_values, dtype=dtype, device=_dev()) if me == rank else None
    return broadcast_tensor((size,), dtype, tensor=t, rank=rank)


def broadcast_int_list(size: int, int_list=None, rank: int = 0, data_parallel: bool = False):
    return broadcast_list(size, torch.int64, list_values=int_list, rank=rank).tolist() if _real() else list(int_list)


def broadcast_float_list(size: int, float_list=None, rank: int = 0, data_parallel: bool = False):
    return broadcast_list(size, torch.float32, list_values=float_list, rank=rank).tolist() if _real() else list(float_list)


def recv_from_prev_pipeline_rank_(recv_buffer=None):
    if not ps.is_pipeline_first_stage():
        dist.recv(recv_buffer, src=ps.get_pipeline_model_parallel_prev_rank())


def send_to_next_pipeline_rank(tensor=None):
    if not ps.is_pipeline_last_stage():
        dist.send(tensor, dst=ps.get_pipeline_model_parallel_next_rank())


def broadcast_from_last_pipeline_stage(size, dtype, tensor=None):
    """The last pipeline stage's tensor on every stage of the pipeline group."""
    if ps.get_pipeline_model_parallel_world_size() == 1:
        return tensor
    buf = tensor.contiguous() if ps.is_pipeline_last_stage() else torch.empty(size, dtype=dtype, device=_dev())
    dist.broadcast(buf, ps.get_pipeline_model_parallel_last_rank(), group=ps.get_pipeline_model_parallel_group())
    return buf


def broadcast_from_last_to_first_pipeline_stage(size, dtype, tensor=None):
    """Last stage -> first stage (other stages get ``None``)."""
    if ps.get_pipeline_model_parallel_world_size() == 1:
        return tensor
    first, last = ps.is_pipeline_first_stage(), ps.is_pipeline_last_stage()
    if not (first or last):
        return None
    if last:
        dist.send(tensor.contiguous(), dst=ps.get_pipeline_model_parallel_first_rank())
        return tensor
    buf = torch.empty(size, dtype=dtype, device=_dev())
    dist.recv(buf, src=ps.get_pipeline_model_parallel_last_rank())
    return buf


def copy_from_last_to_first_pipeline_stage(size, dtype, tensor=None):
    """In place variant: the first stage's ``tensor`` is overwritten with the last stage's."""
    got = broadcast_from_last_to_first_pipeline_stage(size, dtype, tensor)
    if ps.get_pipeline_model_parallel_world_size() > 1 and ps.is_pipeline_first_stage() and got is not None:
        tensor.copy_(got)
