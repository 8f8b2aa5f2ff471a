"""Inference: requests, sampling, KV-cache contexts, scheduler, text-generation controller, engines.

Parity: reference ``megatron/core/inference/**`` -- ``InferenceRequest`` / ``Status``,
``SamplingParams``, ``Scheduler`` (waiting / active / completed pools), ``StaticInferenceContext``
(``key_value_memory_dict``, ``sequence_len_offset``), ``TextGenerationController``
(``generate_all_output_tokens_static_batch`` :388, ``sample_from_logits`` :112, MegaScope
``tik_result`` :582 / ``tik_end`` :639), ``StaticInferenceEngine`` / ``DynamicInferenceEngine``.
"""
from __future__ import annotations

import itertools
import time
from collections import OrderedDict
import dataclasses
from dataclasses import dataclass, field
from enum import Enum
from typing import Dict, List, Optional

import torch

from ..scope import hooks as scope_hooks


class Status(Enum):
    WAITING_IN_QUEUE = 1
    ACTIVE_AND_GENERATING_TOKENS = 2
    ACTIVE_BUT_NOT_GENERATING_TOKENS = 3
    COMPLETED = 4


@dataclass
class SamplingParams:
    temperature: float = 1.0
    top_k: int = 0
    top_p: float = 0.0
    return_log_probs: bool = False
    return_segments: bool = False
    num_tokens_to_generate: int = 30
    top_n_logprobs: int = 0
    add_BOS: bool = False
    # legacy REST API knobs (reference text_generation/generation.py:236-300)
    top_p_decay: float = 0.0                 # top_p is multiplied by this after every generated token ...
    top_p_bound: float = 0.0                 # ... but not below this
    stop_on_eol: bool = False                # stop a row at the first "\n" (or "\n\n") token
    stop_on_double_eol: bool = False         # stop at "\n\n" or two consecutive "\n"
    prevent_newline_after_colon: bool = False
    use_eod_token_for_early_termination: bool = True

    def add_attributes(self, d: dict):
        for k, v in d.items():
            setattr(self, k, v)


@dataclass
class InferenceRequest:
    request_id: str
    prompt: str
    inference_parameters: SamplingParams = None
    prompt_tokens: List[int] = None
    arrival_time: float = 0.0
    status: Status = Status.WAITING_IN_QUEUE
    generated_text: Optional[str] = None
    generated_tokens: Optional[torch.Tensor] = None
    generated_log_probs: Optional[List[float]] = None
    generated_length: int = 0


class StaticInferenceContext:
    """Pre-sized KV cache bookkeeping shared by every attention layer."""

    def __init__(self, max_batch_size: int, max_sequence_length: int):
        self.max_batch_size, self.max_sequence_length = max_batch_size, max_sequence_length
        self.sequence_len_offset = 0
        self.batch_size_offset = 0
        self.key_value_memory_dict: Dict[int, tuple] = {}
        self.decode_mode = False

    def reset(self):
        self.sequence_len_offset = 0
        self.batch_size_offset = 0

    def swap_key_value_dict(self, batch_idx):
        for k, v in list(self.key_value_memory_dict.items()):
            if isinstance(k, tuple) and k[0] == "mamba":           # [conv_state, ssm_state], batch-major
                self.key_value_memory_dict[k] = [t[batch_idx].contiguous() for t in v]
            else:
                kc, vc = v
                self.key_value_memory_dict[k] = (kc[:, batch_idx].contiguous(), vc[:, batch_idx].contiguous())

    def is_static_batching(self):
        return True


InferenceParams = StaticInferenceContext


class Scheduler:
    def __init__(self, max_batch_size: int):
        self.max_batch_size = max_batch_size
        self.active_request_pool: "OrderedDict[str, InferenceRequest]" = OrderedDict()
        self.waiting_request_pool: "OrderedDict[str, InferenceRequest]" = OrderedDict()
        self.completed_request_pool: "OrderedDict[str, InferenceRequest]" = OrderedDict()
        self._ids = itertools.count()

    def add_request(self, prompt: str, prompt_tokens: List[int], inference_parameters: SamplingParams,
                    arrival_time: Optional[float] = None) -> str:
        rid = str(next(self._ids))
        req = InferenceRequest(rid, prompt, inference_parameters, prompt_tokens, arrival_time or time.time())
        if len(self.active_request_pool) < self.max_batch_size:
            req.status = Status.ACTIVE_BUT_NOT_GENERATING_TOKENS
            self.active_request_pool[rid] = req
        else:
            self.waiting_request_pool[rid] = req
        return rid

    def have_requests_pending(self) -> bool:
        return bool(self.active_request_pool) or bool(self.waiting_request_pool)

    def add_earliest_waiting_request_to_active_pool(self):
        if self.waiting_request_pool and len(self.active_request_pool) < self.max_batch_size:
            rid, req = self.waiting_request_pool.popitem(last=False)
            req.status = Status.ACTIVE_BUT_NOT_GENERATING_TOKENS
            self.active_request_pool[rid] = req

    def update_requests_pools(self, result: Optional[Dict[str, InferenceRequest]] = None):
        for rid in list(self.active_request_pool):
            req = self.active_request_pool[rid]
            if req.status == Status.COMPLETED:
                self.completed_request_pool[rid] = self.active_request_pool.pop(rid)
        while self.waiting_request_pool and len(self.active_request_pool) < self.max_batch_size:
            self.add_earliest_waiting_request_to_active_pool()


def sample_from_logits(last_token_logits: torch.Tensor, params: SamplingParams, vocab_size: Optional[int] = None,
                       generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Greedy when ``top_k == 1``; otherwise temperature -> top-k XOR top-p filtering -> multinomial."""
    top_k, top_p, temp = params.top_k, params.top_p, params.temperature
    if not (0.0 <= top_p <= 1.0):
        raise ValueError("top-p should be in [0, 1]")
    if top_k > 0 and top_p > 0.0:
        raise ValueError("cannot set both greedy/top-k and top-p sampling")
    logits = last_token_logits.float()
    if vocab_size:
        logits = logits[..., :vocab_size] if logits.shape[-1] > vocab_size else logits
    if top_k == 1:
        return torch.argmax(logits, dim=-1)
    logits = logits.clone()
    if temp != 1.0:
        logits.div_(temp)
    if top_k > 1:
        if top_k > logits.size(-1):
            raise ValueError("top-k is larger than logit size")
        kth = torch.topk(logits, top_k)[0][..., -1, None]
        logits.masked_fill_(logits < kth, float("-inf"))
    elif top_p > 0.0:
        sorted_logits, sorted_idx = torch.sort(logits, descending=True)
        cum = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
        mask = cum > top_p
        mask[..., 1:] = mask[..., :-1].clone()
        mask[..., 0] = False
        logits.masked_fill_(mask.scatter(1, sorted_idx, mask), float("-inf"))
    probs = logits.softmax(dim=-1)
    return torch.multinomial(probs, num_samples=1, generator=generator).view(-1)


class GPTInferenceWrapper:
    """Drives a GPTModel for prefill + single-token decode steps with a shared KV cache."""

    def __init__(self, model, max_batch_size: int = 8, max_sequence_length: int = 2560, padded_vocab_size=None):
        self.model = model
        m = model
        while hasattr(m, "module"):
            m = m.module
        self.inner = m
        self.max_batch_size, self.max_sequence_length = max_batch_size, max_sequence_length
        self.padded_vocab_size = padded_vocab_size
        self.context: Optional[StaticInferenceContext] = None

    def prep_model_for_inference(self, prompts_tokens: torch.Tensor):
        self.model.eval()
        self.context = StaticInferenceContext(prompts_tokens.shape[0], self.max_sequence_length)

    @torch.no_grad()
    def run_one_forward_step(self, tokens: torch.Tensor, start: int, end: int) -> Optional[torch.Tensor]:
        """tokens [b, S_total]; runs positions [start, end) against the cache; returns logits [b, end-start, V]
        (full vocabulary) on the last pipeline stage and ``None`` elsewhere.  Under pipeline parallelism the hidden
        states hop stage to stage over the p2p channels (reference ``inference/communication_utils.py``); under
        tensor parallelism the vocab-parallel logits are gathered."""
        ctx = self.context
        ctx.sequence_len_offset = start
        ids = tokens[:, start:end]
        pos = torch.arange(start, end, device=tokens.device).unsqueeze(0).expand_as(ids)
        out = pipeline_forward(self.model, self.inner, ids, pos, ctx, (end - start, ids.shape[0]))
        ctx.sequence_len_offset = end
        return self._full_vocab(out)

    def _full_vocab(self, logits):
        return gather_full_vocab(logits, self.padded_vocab_size or getattr(self.inner, "vocab_size", None))


def pipeline_forward(model, inner, ids, pos, ctx, seq_batch):
    """One forward of ``ids`` through every pipeline stage: the hidden states hop stage to stage over the p2p channels.
    Returns the model output on the last stage and ``None`` on the others.  ``seq_batch`` = (tokens, batch) of the
    ``[s, b, h]`` activation that crosses the stage boundaries."""
    from ..parallel import p2p, state as ps

    pp = ps.get_pipeline_model_parallel_world_size() if ps.model_parallel_is_initialized() else 1
    if pp == 1:
        return model(ids, pos, None, inference_context=ctx)
    cfg = inner.config
    if not ps.is_pipeline_first_stage():
        h = p2p.recv((seq_batch[0], seq_batch[1], cfg.hidden_size), cfg.pipeline_dtype or cfg.params_dtype,
                     ps.get_pipeline_model_parallel_prev_rank(), "fwd", "recv-forward")
        inner.set_input_tensor(h)
    out = model(ids, pos, None, inference_context=ctx)
    if not ps.is_pipeline_last_stage():
        p2p.send(out, ps.get_pipeline_model_parallel_next_rank(), "fwd", "send-forward")
        p2p.flush()
        return None
    return out


def gather_full_vocab(logits, full_vocab: Optional[int]):
    """Vocab-parallel logits -> full vocabulary (no-op without TP or when the model already gathered)."""
    from ..parallel import state as ps

    if logits is None or not ps.model_parallel_is_initialized() or ps.get_tensor_model_parallel_world_size() == 1:
        return logits
    if full_vocab is not None and logits.shape[-1] >= full_vocab:
        return logits
    from ..parallel.mappings import gather_from_tensor_model_parallel_region

    return gather_from_tensor_model_parallel_region(logits)


def sync_sampled_tokens(new: Optional[torch.Tensor], batch: int, device) -> torch.Tensor:
    """Every model-parallel rank continues with the SAME next tokens: the last pipeline stage samples, then the
    tokens go to the other stages (PP group) and, inside each stage, from TP rank 0 to its peers."""
    from ..parallel import state as ps

    if not ps.model_parallel_is_initialized():
        return new
    import torch.distributed as dist

    if new is None:
        new = torch.zeros(batch, dtype=torch.long, device=device)
    if ps.get_pipeline_model_parallel_world_size() > 1:
        dist.broadcast(new, ps.get_pipeline_model_parallel_last_rank(), group=ps.get_pipeline_model_parallel_group())
    if ps.get_tensor_model_parallel_world_size() > 1:
        dist.broadcast(new, ps.get_tensor_model_parallel_src_rank(), group=ps.get_tensor_model_parallel_group())
    return new


class VLMInferenceWrapper(GPTInferenceWrapper):
    """LLaVA-style generation (reference ``inference/model_inference_wrappers/multimodal/vlm_inference_wrapper.py``):
    the prefill runs the vision tower and splices the image embeddings in, which makes the cached sequence longer
    than the token sequence by ``image_tokens_count``; decode steps feed text tokens only and shift the cache /
    rotary offset by that amount (embedding position ids stay those of the text stream)."""

    def __init__(self, model, max_batch_size: int = 8, max_sequence_length: int = 2560, padded_vocab_size=None):
        super().__init__(model, max_batch_size, max_sequence_length, padded_vocab_size)
        self.images = self.num_image_tiles = None
        self._extra = 0

    def set_images(self, images: torch.Tensor, num_image_tiles: Optional[torch.Tensor] = None):
        self.images, self.num_image_tiles = images, num_image_tiles

    def prep_model_for_inference(self, prompts_tokens: torch.Tensor):
        super().prep_model_for_inference(prompts_tokens)
        self._extra = 0

    @torch.no_grad()
    def run_one_forward_step(self, tokens: torch.Tensor, start: int, end: int) -> torch.Tensor:
        ctx, m = self.context, self.inner
        ids = tokens[:, start:end]
        pos = torch.arange(start, end, device=tokens.device).unsqueeze(0).expand_as(ids)
        if start == 0:
            n_img = (ids == m.image_token_index).sum(dim=1)
            if int((tokens == m.image_token_index).sum()) != int(n_img.sum()):
                raise ValueError("every image token must lie inside the shortest prompt of the batch (prefill window)")
            if len(set(n_img.tolist())) > 1:
                raise ValueError("all prompts of a batch must carry the same number of images")
            tiles = self.num_image_tiles
            per_img = m.img_seq_len if tiles is None else None
            self._extra = int(n_img[0]) * (m.img_seq_len - 1) if per_img else \
                int(tiles.sum()) // ids.shape[0] * m.img_seq_len - int(n_img[0])
            ctx.sequence_len_offset = 0
            logits = self.model(self.images, ids, pos, None, inference_params=ctx, num_image_tiles=tiles)
            ctx.key_value_memory_dict["image_tokens_count"] = self._extra
        else:
            ctx.sequence_len_offset = start + self._extra
            logits = self.model(None, ids, pos, None, inference_params=ctx)
        ctx.sequence_len_offset = end + self._extra
        return logits


class T5InferenceWrapper(GPTInferenceWrapper):
    """Encoder-decoder generation (reference ``model_inference_wrappers/t5/t5_inference_wrapper.py``): the prompt is the
    ENCODER input; it is encoded once per batch and cached, and the decoder -- started from ``decoder_start_token`` --
    is re-run on the tokens generated so far at every step (cross attention to the cached encoder states, causal self
    attention), which needs no decoder KV cache and keeps the model code untouched."""

    def __init__(self, model, max_batch_size: int = 8, max_sequence_length: int = 512, padded_vocab_size=None,
                 decoder_start_token: int = 0, pad_token: int = 0):
        super().__init__(model, max_batch_size, max_sequence_length, padded_vocab_size)
        self.decoder_start_token, self.pad_token = decoder_start_token, pad_token
        self._enc = self._enc_mask = None

    def prep_model_for_inference(self, prompts_tokens: torch.Tensor):
        super().prep_model_for_inference(prompts_tokens)
        self._enc = self._enc_mask = None

    @torch.no_grad()
    def encode(self, encoder_tokens: torch.Tensor):
        pad = (encoder_tokens != self.pad_token).float()
        self._enc_mask = pad
        self._enc = self.model(encoder_tokens, None, pad.unsqueeze(1) * pad.unsqueeze(2), None, None,
                               output_encoder_hidden_only=True)
        return self._enc

    @torch.no_grad()
    def decode_step(self, decoder_tokens: torch.Tensor) -> torch.Tensor:
        """``decoder_tokens [b, t]`` (starting with ``decoder_start_token``) -> next-token logits ``[b, V]``."""
        b, t = decoder_tokens.shape
        dm = torch.tril(torch.ones(t, t, device=decoder_tokens.device)).unsqueeze(0).expand(b, -1, -1)
        edm = self._enc_mask.unsqueeze(1).expand(b, t, -1)
        logits = self.model(None, decoder_tokens, None, dm, edm, encoder_hidden_states=self._enc)
        return gather_full_vocab(logits, self.padded_vocab_size)[:, -1]

    @torch.no_grad()
    def generate(self, encoder_tokens: torch.Tensor, params: "SamplingParams", eos_token: Optional[int] = None):
        """Greedy / sampled decoding for a batch of encoder inputs; returns ``[b, <= num_tokens_to_generate]`` token ids
        (rows stop at ``eos_token``, the remainder is padded with it)."""
        self.model.eval()
        self.encode(encoder_tokens)
        b, dev = encoder_tokens.shape[0], encoder_tokens.device
        dec = torch.full((b, 1), self.decoder_start_token, dtype=torch.long, device=dev)
        done = torch.zeros(b, dtype=torch.bool, device=dev)
        for _ in range(params.num_tokens_to_generate):
            nxt = sync_sampled_tokens(sample_from_logits(self.decode_step(dec), params), b, dev)
            if eos_token is not None:
                nxt = torch.where(done, torch.full_like(nxt, eos_token), nxt)
                done |= nxt == eos_token
            dec = torch.cat([dec, nxt.unsqueeze(1)], dim=1)
            if eos_token is not None and bool(done.all()):
                break
        return dec[:, 1:]


class TextGenerationController:
    def __init__(self, inference_wrapped_model: GPTInferenceWrapper, tokenizer):
        self.inference_wrapped_model = inference_wrapped_model
        self.tokenizer = tokenizer

    def tokenize_prompt(self, prompt: str, add_BOS: bool = False) -> List[int]:
        toks = self.tokenizer.tokenize(prompt)
        if add_BOS and getattr(self.tokenizer, "bos", None) is not None:
            toks = [self.tokenizer.bos] + toks
        return toks

    def detokenize_generations(self, tokens: List[int]) -> str:
        return self.tokenizer.detokenize(tokens)

    def sample_from_logits(self, logits, params, vocab_size=None, generator=None):
        return sample_from_logits(logits, params, vocab_size, generator)

    @torch.no_grad()
    def generate_all_output_tokens_static_batch(self, active_requests: "OrderedDict[str, InferenceRequest]"):
        reqs = list(active_requests.values())
        params = reqs[0].inference_parameters
        dev = next(self.inference_wrapped_model.model.parameters()).device
        lens = [len(r.prompt_tokens) for r in reqs]
        min_len, max_len = min(lens), max(lens)
        n_gen = params.num_tokens_to_generate
        total = min(max_len + n_gen, self.inference_wrapped_model.max_sequence_length)
        eod = self.tokenizer.eod
        tokens = torch.full((len(reqs), total), eod, dtype=torch.long, device=dev)
        for i, r in enumerate(reqs):
            tokens[i, :lens[i]] = torch.tensor(r.prompt_tokens, dtype=torch.long, device=dev)
        lens_t = torch.tensor(lens, device=dev)
        self.inference_wrapped_model.prep_model_for_inference(tokens)
        done = torch.zeros(len(reqs), dtype=torch.bool, device=dev)
        gen_len = torch.zeros(len(reqs), dtype=torch.long, device=dev)
        logprobs: List[List[float]] = [[] for _ in reqs]
        vocab = getattr(self.tokenizer, "vocab_size", None)
        start = 0
        top_p = params.top_p
        eol = eol2 = colon = None
        if params.stop_on_eol or params.stop_on_double_eol or params.prevent_newline_after_colon:
            # tokenizer dependent, as in the reference (which hard-codes GPT-2's 198 / 628): the ids of "\n", "\n\n" and ":"
            def one(text):
                ids = self.tokenizer.tokenize(text)
                return int(ids[0]) if len(ids) == 1 else None

            eol, eol2, colon = one("\n"), one("\n\n"), one(":")
        for pos in range(min_len, total):
            logits = self.inference_wrapped_model.run_one_forward_step(tokens, start, pos)
            last = new = None
            if logits is not None:                      # last pipeline stage
                last = logits[:, -1, :]
                if params.prevent_newline_after_colon and colon is not None and eol is not None:
                    last = last.clone()
                    last[tokens[:, pos - 1] == colon, eol] = -1e10
                step = params if top_p == params.top_p else dataclasses.replace(params, top_p=top_p)
                new = self.sample_from_logits(last, step, vocab)
                if top_p > 0.0 and params.top_p_decay > 0.0:
                    top_p = top_p * params.top_p_decay
                    if params.top_p_bound > 0.0:
                        top_p = max(top_p, params.top_p_bound)
                scope_hooks.tik_result(last, new)
            new = sync_sampled_tokens(new, len(reqs), dev)
            started = lens_t <= pos                     # rows still consuming their prompt keep it
            tokens[started, pos] = new[started]
            if params.return_log_probs and last is not None:
                lp = torch.log_softmax(last.float(), dim=-1)
                for i in range(len(reqs)):
                    if started[i] and not done[i]:
                        logprobs[i].append(float(lp[i, tokens[i, pos]]))
            never = torch.zeros_like(started)
            is_eol2 = (new == eol2) if eol2 is not None else never
            is_eol = (new == eol) if eol is not None else never
            if params.stop_on_double_eol:
                hit = is_eol2 | (is_eol & (tokens[:, pos - 1] == eol))
            elif params.stop_on_eol:
                hit = is_eol2 | is_eol
            elif params.use_eod_token_for_early_termination:
                hit = new == eod
            else:
                hit = never
            hit = started & hit & ~done
            gen_len += (started & ~done).long()
            done |= hit | (gen_len >= n_gen)            # a short prompt in a ragged batch still stops after n_gen tokens
            start = pos
            if bool(done.all()):
                break
        scope_hooks.tik_end()
        for i, r in enumerate(reqs):
            n = int(gen_len[i])
            out = tokens[i, lens[i]:lens[i] + n]
            r.generated_tokens = out
            r.generated_length = n
            r.generated_log_probs = logprobs[i] if params.return_log_probs else None
            text = out.tolist()
            if text and text[-1] == eod:
                text = text[:-1]
            r.generated_text = self.detokenize_generations(text)
            r.status = Status.COMPLETED
        return active_requests


class StaticInferenceEngine:
    def __init__(self, text_generation_controller: TextGenerationController, max_batch_size: int = 8,
                 random_seed: Optional[int] = None):
        self.controller = text_generation_controller
        self.scheduler = Scheduler(max_batch_size)
        self.random_seed = random_seed

    def generate(self, prompts: List[str], sampling_params: Optional[SamplingParams] = None,
                 add_BOS: bool = False) -> List[InferenceRequest]:
        sampling_params = sampling_params or SamplingParams()
        if self.random_seed is not None:
            torch.manual_seed(self.random_seed)
        ids = []
        for p in prompts:
            ids.append(self.scheduler.add_request(p, self.controller.tokenize_prompt(p, add_BOS), sampling_params))
        self.run_engine()
        return [self.scheduler.completed_request_pool[i] for i in ids]

    def run_engine(self):
        while self.scheduler.have_requests_pending():
            active = self.scheduler.active_request_pool
            self.controller.generate_all_output_tokens_static_batch(active)
            self.scheduler.update_requests_pools(active)


# -- base-class and config names of the reference's ``core/inference`` tree ------------------------------------------------------
AbstractModelInferenceWrapper = GPTInferenceWrapper          # the GPT wrapper is the base of the T5 / VLM ones here
BaseInferenceContext = StaticInferenceContext


class AbstractEngine:
    """Interface of an inference engine: ``generate(prompts, sampling_params) -> List[InferenceRequest]``."""

    def generate(self, prompts, sampling_params=None, **kw):
        raise NotImplementedError


@dataclass
class InferenceWrapperConfig:
    """Reference ``inference_wrapper_config.py``: the knobs a model inference wrapper is built from."""
    hidden_size: int = 0
    params_dtype: torch.dtype = torch.float32
    inference_batch_times_seqlen_threshold: int = -1
    padded_vocab_size: int = 0
    inference_max_requests: int = 8
    inference_max_seq_length: int = 2560
    fp32_residual_connection: bool = False
    nccl_all_reduce_for_prefill: bool = False

    def add_attributes(self, attribute_value_pair: dict):
        for k, v in attribute_value_pair.items():
            setattr(self, k, v)


class Counter:
    """Monotonic request-id source (reference ``inference/utils.py``)."""

    def __init__(self, start: int = 0) -> None:
        self.counter = start

    def __next__(self) -> int:
        i = self.counter
        self.counter += 1
        return i

    def reset(self) -> None:
        self.counter = 0


@dataclass
class VLMInferenceRequest(InferenceRequest):
    """An ``InferenceRequest`` that also carries the images of its prompt."""
    num_img_embeddings_per_tile: int = 0
    imgs: Optional[torch.Tensor] = None
    num_tiles: Optional[torch.Tensor] = None
    decoder_seq_length: int = 0


class EncoderDecoderTextGenerationController(TextGenerationController):
    """T5-style generation: the wrapper (``T5InferenceWrapper``) encodes once and decodes step by step."""


class VLMTextGenerationController(TextGenerationController):
    """LLaVA-style generation: the wrapper (``VLMInferenceWrapper``) splices the image embeddings in at prefill."""
