"""Function-style text generation (reference ``megatron/inference/text_generation/{api,generation,tokenization}.py``).

The reference threads an ``inference_context`` and a ``forward_step`` class through these calls; here the model is wrapped in
``GPTInferenceWrapper`` + ``StaticInferenceEngine`` (KV cache, pipeline / tensor-parallel token synchronisation) on first use
and the wrapper is cached on the model, so callers keep the one-line API:

    texts, segments, logprobs, tokens = generate_and_post_process(model, prompts=["Hello"], tokens_to_generate=32, top_k_sampling=1)
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from ..beam_search import beam_search as _beam_search
from ..core import GPTInferenceWrapper, SamplingParams, StaticInferenceEngine, TextGenerationController


def _tokenizer(tokenizer=None):
    if tokenizer is not None:
        return tokenizer
    from ...training.global_vars import get_tokenizer

    return get_tokenizer()


def _engine(model, tokenizer, max_batch: int, max_seq: Optional[int]) -> StaticInferenceEngine:
    inner = model
    while hasattr(inner, "module"):
        inner = inner.module
    cache = getattr(inner, "_b200_generation_engine", None)
    if cache is not None and cache[1] >= max_batch and cache[2] is tokenizer:
        return cache[0]
    if max_seq is None:
        max_seq = getattr(inner, "max_sequence_length", None) or getattr(getattr(inner, "config", None), "seq_length", None) or 4096
    pv = getattr(inner, "vocab_size", None)
    eng = StaticInferenceEngine(TextGenerationController(GPTInferenceWrapper(model, max_batch, max_seq, pv), tokenizer), max_batch)
    inner._b200_generation_engine = (eng, max_batch, tokenizer)
    return eng


def _check_token_budget(tok, prompts, tokens_to_generate: int, add_BOS: bool) -> None:
    """``--max-tokens-to-oom`` (reference generation.py:170): refuse a batch whose padded size would not fit."""
    try:
        from ...training.global_vars import get_args

        limit = getattr(get_args(), "max_tokens_to_oom", None)
    except Exception:  # noqa: BLE001  (library use without the global args)
        limit = None
    if not limit:
        return
    longest = max(len(tok.tokenize(p)) + int(bool(add_BOS)) for p in prompts) + int(tokens_to_generate)
    if longest * len(prompts) > limit:
        raise ValueError(f"Too many tokens.  {longest * len(prompts)} is greater than {limit}")


def generate(model, inference_context=None, forward_step=None, prompts: Optional[Sequence[str]] = None, tokens_to_generate: int = 0,
             return_output_log_probs: bool = False, top_k_sampling: int = 0, top_p_sampling: float = 0.0, top_p_decay: float = 0.0,
             top_p_bound: float = 0.0, temperature: float = 1.0, add_BOS: bool = False, use_eod_token_for_early_termination: bool = True,
             stop_on_double_eol: bool = False, stop_on_eol: bool = False, prevent_newline_after_colon: bool = False,
             random_seed: int = -1, data_parallel: bool = False, return_topk_logprobs: int = 0, tokenizer=None,
             max_sequence_length: Optional[int] = None):
    """Returns ``(tokens [b, max_len] padded with eod, lengths [b], output_log_probs | None, None)``."""
    if not prompts:
        raise ValueError("prompts is empty")
    tok = _tokenizer(tokenizer)
    _check_token_budget(tok, prompts, tokens_to_generate, add_BOS)
    if random_seed is not None and random_seed >= 0:
        torch.manual_seed(random_seed)
    eng = _engine(model, tok, max(len(prompts), 1), max_sequence_length)
    sp = SamplingParams(temperature=temperature, top_k=top_k_sampling, top_p=top_p_sampling, return_log_probs=return_output_log_probs,
                        num_tokens_to_generate=tokens_to_generate, add_BOS=add_BOS, top_p_decay=top_p_decay, top_p_bound=top_p_bound,
                        stop_on_eol=stop_on_eol, stop_on_double_eol=stop_on_double_eol,
                        prevent_newline_after_colon=prevent_newline_after_colon,
                        use_eod_token_for_early_termination=use_eod_token_for_early_termination)
    res = eng.generate(list(prompts), sp, add_BOS=add_BOS)
    rows = [list(r.prompt_tokens) + [int(t) for t in r.generated_tokens] for r in res]
    lengths = torch.tensor([len(r) for r in rows])
    tokens = torch.full((len(rows), int(lengths.max())), tok.eod, dtype=torch.long)
    for i, r in enumerate(rows):
        tokens[i, :len(r)] = torch.tensor(r)
    logp = [list(r.generated_log_probs or []) for r in res] if return_output_log_probs else None
    return tokens, lengths, logp, None


def generate_and_post_process(model, inference_context=None, forward_step=None, prompts=None, tokens_to_generate=0,
                              return_output_log_probs=False, top_k_sampling=0, top_p_sampling=0.0, top_p_decay=0.0, top_p_bound=0.0,
                              temperature=1.0, add_BOS=False, use_eod_token_for_early_termination=True, stop_on_double_eol=False,
                              stop_on_eol=False, prevent_newline_after_colon=False, random_seed=-1, detokenize_segments=True,
                              data_parallel=False, return_topk_logprobs=0, tokenizer=None, max_sequence_length=None):
    """Returns ``(texts, segments, output_log_probs, tokens)`` as python lists: ``texts`` = prompt + continuation,
    ``segments`` = the per-token strings of each row (``detokenize_segments``), ``tokens`` = token id lists."""
    tok = _tokenizer(tokenizer)
    tokens, lengths, logp, _ = generate(model, inference_context, forward_step, prompts, tokens_to_generate, return_output_log_probs,
                                        top_k_sampling, top_p_sampling, top_p_decay, top_p_bound, temperature, add_BOS,
                                        use_eod_token_for_early_termination, stop_on_double_eol, stop_on_eol, prevent_newline_after_colon,
                                        random_seed, data_parallel, return_topk_logprobs, tok, max_sequence_length)
    rows = [tokens[i, : int(lengths[i])].tolist() for i in range(tokens.shape[0])]
    texts = [tok.detokenize([t for t in r if t != tok.eod]) for r in rows]
    segments = [[tok.detokenize([t]) for t in r] for r in rows] if detokenize_segments else None
    return texts, segments, logp, rows


def beam_search(model, forward_step=None, prompts=None, tokens_to_generate: int = 0, beam_size: int = 0, add_BOS: bool = False,
                stop_token: Optional[int] = None, num_return_gen: int = 1, length_penalty: float = 1.0,
                prevent_newline_after_colon: bool = False, tokenizer=None, max_sequence_length: Optional[int] = None):
    """Single prompt (as in the reference).  Returns ``(tokens [num_return_gen, len], scores [num_return_gen])``."""
    if not prompts or len(prompts) != 1:
        raise ValueError("beam search takes exactly one prompt")
    tok = _tokenizer(tokenizer)
    eng = _engine(model, tok, max(beam_size, 1), max_sequence_length)
    ids = eng.controller.tokenize_prompt(prompts[0], add_BOS)
    return _beam_search(eng.controller.inference_wrapped_model, ids, beam_size, tokens_to_generate,
                        tok.eod if stop_token is None else stop_token, length_penalty, num_return_gen)


def beam_search_and_post_process(model, forward_step=None, prompts=None, tokens_to_generate=0, beam_size=0, add_BOS=False,
                                 stop_token=None, num_return_gen=1, length_penalty=1.0, prevent_newline_after_colon=False,
                                 detokenize_segments=True, tokenizer=None, max_sequence_length=None):
    """Returns ``(texts, segments, scores)``."""
    tok = _tokenizer(tokenizer)
    hyps, scores = beam_search(model, forward_step, prompts, tokens_to_generate, beam_size, add_BOS, stop_token, num_return_gen,
                               length_penalty, prevent_newline_after_colon, tok, max_sequence_length)
    rows = [[int(t) for t in h] for h in hyps]
    texts = [tok.detokenize([t for t in r if t != tok.eod]) for r in rows]
    segments = [[tok.detokenize([t]) for t in r] for r in rows] if detokenize_segments else None
    return texts, segments, [float(s) for s in scores]


# -- token-level entry points of the reference's ``generation.py`` --------------------------------------------------------------------
def _rows(tokens: torch.Tensor, lengths: torch.Tensor) -> List[List[int]]:
    return [tokens[i, : int(lengths[i])].tolist() for i in range(tokens.shape[0])]


def generate_tokens_probs_and_return_on_first_stage(model, forward_step=None, tokens=None, lengths=None, return_output_log_probs=False,
                                                    top_k=0, top_p=0.0, top_p_decay=0.0, top_p_bound=0.0, temperature=1.0,
                                                    use_eod_token_for_early_termination=True, stop_on_double_eol=False,
                                                    stop_on_eol=False, prevent_newline_after_colon=True, tokenizer=None,
                                                    max_sequence_length=None, inference_context=None):
    """Generation from already tokenised, right-padded prompts (``tokens [b, max_prompt + tokens_to_generate]``, ``lengths [b]`` =
    prompt lengths): the continuation is written into ``tokens`` in place.  Returns ``(tokens, generated sequence lengths,
    output_log_probs [b, max_len - 1] | None, None)`` like the reference."""
    tok = _tokenizer(tokenizer)
    b, total = tokens.shape
    n_new = total - int(lengths.min())
    eng = _engine(model, tok, b, max_sequence_length or total)
    sp = SamplingParams(temperature=temperature, top_k=top_k, top_p=top_p, return_log_probs=return_output_log_probs,
                        num_tokens_to_generate=n_new, top_p_decay=top_p_decay, top_p_bound=top_p_bound, stop_on_eol=stop_on_eol,
                        stop_on_double_eol=stop_on_double_eol, prevent_newline_after_colon=prevent_newline_after_colon,
                        use_eod_token_for_early_termination=use_eod_token_for_early_termination)
    ids = [eng.scheduler.add_request(tok.detokenize(r), r, sp) for r in _rows(tokens, lengths)]
    eng.run_engine()
    res = [eng.scheduler.completed_request_pool[i] for i in ids]
    out_lengths = lengths.clone()
    logp = torch.zeros(b, total - 1, dtype=torch.float32, device=tokens.device) if return_output_log_probs else None
    for i, r in enumerate(res):
        n0 = int(lengths[i])
        gen = [int(t) for t in r.generated_tokens][: total - n0]
        tokens[i, n0:n0 + len(gen)] = torch.tensor(gen, dtype=tokens.dtype, device=tokens.device)
        out_lengths[i] = n0 + len(gen)
        if logp is not None and r.generated_log_probs:
            lp = list(r.generated_log_probs)[: len(gen)]
            logp[i, n0 - 1:n0 - 1 + len(lp)] = torch.tensor(lp, device=tokens.device)
    return tokens, out_lengths, logp, None


@torch.no_grad()
def score_and_return_on_first_stage(model, tokens: torch.Tensor, lengths: torch.Tensor, **_):
    """Log-probability of every token given its prefix under the model: ``(tokens, lengths, log_probs [b, s - 1], None)``
    (position ``j`` scores ``tokens[:, j + 1]``).  One full forward pass; vocabulary-parallel logits are gathered."""
    from ..core import gather_full_vocab

    inner = model
    while hasattr(inner, "module"):
        inner = inner.module
    was_training = inner.training
    inner.eval()
    try:
        pos = torch.arange(tokens.shape[1], device=tokens.device).unsqueeze(0).expand_as(tokens)
        logits = model(tokens, pos, None)                                   # [b, s, V / tp]
        logits = gather_full_vocab(logits, getattr(inner, "vocab_size", None))
        lp = torch.log_softmax(logits.float(), dim=-1)
        out = torch.gather(lp[:, :-1], 2, tokens[:, 1:].unsqueeze(2)).squeeze(2)
    finally:
        inner.train(was_training)
    return tokens, lengths, out, None


def beam_search_and_return_on_first_stage(model, forward_step=None, tokens=None, lengths=None, beam_size: int = 1, stop_token=None,
                                          num_return_gen: int = 1, length_penalty: float = 1.0, prevent_newline_after_colon: bool = True,
                                          tokenizer=None, max_sequence_length=None, inference_context=None):
    """Beam search from one tokenised prompt: ``(tokens [num_return_gen, len], scores [num_return_gen])``."""
    if tokens.shape[0] != 1:
        raise ValueError("beam search takes exactly one prompt")
    tok = _tokenizer(tokenizer)
    n0 = int(lengths[0])
    eng = _engine(model, tok, max(beam_size, 1), max_sequence_length or tokens.shape[1])
    return _beam_search(eng.controller.inference_wrapped_model, tokens[0, :n0].tolist(), beam_size, tokens.shape[1] - n0,
                        tok.eod if stop_token is None else stop_token, length_penalty, num_return_gen)
