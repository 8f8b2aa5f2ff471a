"""Dynamic (continuous) batching with a paged KV cache.

Parity: reference ``megatron/core/inference/contexts/dynamic_context.py`` (``DynamicInferenceContext``: chunked
KV memory, per-request block tables, token-granularity batch) and ``engines/dynamic_engine.py``
(``DynamicInferenceEngine.step``: admit -> forward the packed token batch -> sample -> retire / evict).

Design here: every step runs ONE forward over a flat token batch ``[T, 1, h]`` that mixes prefill chunks
(chunked prefill under a per-step token budget) and single decode tokens of all running requests.  The KV
cache is one ``[num_blocks * block_size, groups, head_dim]`` tensor per layer; a request owns a list of
blocks, a token at position ``p`` lives in slot ``table[p // bs] * bs + p % bs``.  Attention gathers the
request's slots (padded over requests), so admission, retirement and preemption never move KV data.
"""
from __future__ import annotations

import os
import itertools
import queue
import threading
import time
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from .core import (InferenceRequest, SamplingParams, Status, gather_full_vocab, pipeline_forward, sample_from_logits,
                   sync_sampled_tokens)


class BlockAllocator:
    def __init__(self, num_blocks: int):
        self.num_blocks = num_blocks
        self._free = list(range(num_blocks - 1, -1, -1))

    @property
    def free(self) -> int:
        return len(self._free)

    def alloc(self, n: int) -> List[int]:
        if n > len(self._free):
            raise MemoryError(f"KV cache exhausted: want {n} blocks, {len(self._free)} free")
        return [self._free.pop() for _ in range(n)]

    def release(self, blocks: List[int]):
        self._free.extend(reversed(blocks))


@dataclass
class _Seq:
    blocks: List[int] = field(default_factory=list)
    kv_len: int = 0            # tokens whose K/V are in the cache


class DynamicInferenceContext:
    def __init__(self, max_sequence_length: int, num_blocks: int, block_size: int = 16,
                 max_tokens_per_step: int = 256, max_requests: int = 64):
        self.max_sequence_length = max_sequence_length
        self.block_size, self.num_blocks = block_size, num_blocks
        self.max_tokens_per_step, self.max_requests = max_tokens_per_step, max_requests
        self.allocator = BlockAllocator(num_blocks)
        self.seqs: Dict[str, _Seq] = {}
        self.kv: Dict[int, tuple] = {}                 # layer -> (K, V) [num_blocks*bs, ng, hn]
        self.use_paged_kernel = os.environ.get("B200_PAGED_ATTN", "1") == "1"     # 0: gather + einsum reference path
        self.kernel_calls = 0
        # per-step plan (set by begin_step)
        self.positions: Optional[torch.Tensor] = None
        self._plan: List[tuple] = []                   # (rid, q_start, n_new, kv_before)
        self._write_slots = self._q_index = self._q_valid = self._kv_slots = self._mask = None
        # attributes the static-path code reads
        self.sequence_len_offset = 0
        self.batch_size_offset = 0
        self.max_batch_size = max_requests

    def is_static_batching(self) -> bool:
        return False

    # -- memory -----------------------------------------------------------------
    def blocks_needed(self, rid: Optional[str], n_new: int) -> int:
        s = self.seqs.get(rid) if rid is not None else None
        have = len(s.blocks) if s else 0
        total = (s.kv_len if s else 0) + n_new
        return max(0, -(-total // self.block_size) - have)

    def add_request(self, rid: str):
        self.seqs[rid] = _Seq()

    def release(self, rid: str):
        s = self.seqs.pop(rid, None)
        if s:
            self.allocator.release(s.blocks)

    def _slot(self, s: _Seq, pos: torch.Tensor) -> torch.Tensor:
        table = torch.tensor(s.blocks, dtype=torch.long)
        return table[pos // self.block_size] * self.block_size + pos % self.block_size

    # -- step plan ---------------------------------------------------------------
    def begin_step(self, work: List[tuple], device) -> torch.Tensor:
        """``work = [(rid, n_new)]``: reserve blocks, build the gather / scatter indices and the mask of the
        packed batch.  Returns the absolute positions ``[T]`` of the new tokens."""
        plan, pos_all, slots_all = [], [], []
        q0 = 0
        for rid, n_new in work:
            s = self.seqs[rid]
            if s.kv_len + n_new > self.max_sequence_length:
                raise ValueError(f"request {rid} exceeds max_sequence_length={self.max_sequence_length}")
            s.blocks += self.allocator.alloc(self.blocks_needed(rid, n_new))
            pos = torch.arange(s.kv_len, s.kv_len + n_new)
            pos_all.append(pos)
            slots_all.append(self._slot(s, pos))
            plan.append((rid, q0, n_new, s.kv_len))
            q0 += n_new
        R = len(plan)
        n_max = max(p[2] for p in plan)
        l_max = max(p[2] + p[3] for p in plan)
        q_index = torch.zeros(R, n_max, dtype=torch.long)
        q_valid = torch.zeros(R, n_max, dtype=torch.bool)
        kv_slots = torch.zeros(R, l_max, dtype=torch.long)
        mask = torch.zeros(R, n_max, l_max, dtype=torch.bool)        # True = may attend
        for i, (rid, qs, n, before) in enumerate(plan):
            q_index[i, :n] = torch.arange(qs, qs + n)
            q_valid[i, :n] = True
            L = before + n
            kv_slots[i, :L] = self._slot(self.seqs[rid], torch.arange(L))
            qpos = torch.arange(before, L).unsqueeze(1)
            mask[i, :n, :L] = torch.arange(L).unsqueeze(0) <= qpos
            mask[i, n:, 0] = True                                      # padded query rows: keep softmax finite
        self._plan = plan
        # paged-attention kernel inputs: the request row and the number of visible keys of every packed query token, and
        # the block tables of this step's requests
        max_blocks = max(len(self.seqs[p[0]].blocks) for p in plan)
        table = torch.zeros(R, max_blocks, dtype=torch.int32)
        tok_req, tok_len = [], []
        for i, (rid, qs, n, before) in enumerate(plan):
            blocks = self.seqs[rid].blocks
            table[i, :len(blocks)] = torch.tensor(blocks, dtype=torch.int32)
            tok_req += [i] * n
            tok_len += list(range(before + 1, before + n + 1))
        self._block_table = table.to(device)
        self._tok_req = torch.tensor(tok_req, dtype=torch.int32).to(device)
        self._tok_len = torch.tensor(tok_len, dtype=torch.int32).to(device)
        self._max_len = l_max
        self.positions = torch.cat(pos_all).to(device)
        self._write_slots = torch.cat(slots_all).to(device)
        self._q_index, self._q_valid = q_index.to(device), q_valid.to(device)
        self._kv_slots, self._mask = kv_slots.to(device), mask.to(device)
        return self.positions

    def end_step(self):
        for rid, _, n, _ in self._plan:
            self.seqs[rid].kv_len += n
        self._plan = []

    def last_token_index(self) -> Dict[str, int]:
        return {rid: qs + n - 1 for rid, qs, n, _ in self._plan}

    # -- attention over the paged cache --------------------------------------------
    def attend(self, layer: int, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float) -> torch.Tensor:
        """``q [T, np, hn]``, new ``k / v [T, ng, hn(v)]`` -> ``[T, np, hn_v]``; writes k / v into their pages first."""
        if layer not in self.kv:
            n_slots = self.num_blocks * self.block_size
            self.kv[layer] = (torch.zeros(n_slots, *k.shape[1:], dtype=k.dtype, device=k.device),
                              torch.zeros(n_slots, *v.shape[1:], dtype=v.dtype, device=v.device))
        kc, vc = self.kv[layer]
        kc[self._write_slots] = k
        vc[self._write_slots] = v
        np_, ng = q.shape[1], k.shape[1]
        if q.is_cuda and self.use_paged_kernel and k.shape[-1] == v.shape[-1] == q.shape[-1] and q.dtype in (torch.bfloat16, torch.float16):
            from .. import ops

            ext = ops.ext()
            if ext.paged_attention_supported(q.shape[-1], np_ // ng):
                # flash-decoding split of long contexts when the (token, KV head) grid alone cannot fill the GPU
                ctas = q.shape[0] * ng
                splits = 1 if ctas >= 296 else max(1, min(16, self._max_len // 512, 296 // max(1, ctas)))
                self.kernel_calls += 1
                return ext.paged_attention(q.contiguous(), kc, vc, self._block_table, self._tok_req, self._tok_len,
                                           self.block_size, float(scale), splits)
        qp = q[self._q_index]                                        # [R, n, np, hn]
        kp, vp = kc[self._kv_slots], vc[self._kv_slots]               # [R, L, ng, hn]
        if np_ != ng:
            kp = kp.repeat_interleave(np_ // ng, dim=2)
            vp = vp.repeat_interleave(np_ // ng, dim=2)
        s = torch.einsum("rnhd,rlhd->rhnl", qp.float(), kp.float()) * scale
        s = s.masked_fill(~self._mask.unsqueeze(1), float("-inf"))
        o = torch.einsum("rhnl,rlhd->rnhd", torch.softmax(s, dim=-1), vp.float()).to(q.dtype)
        out = torch.empty(q.shape[0], np_, vp.shape[-1], dtype=q.dtype, device=q.device)
        out[self._q_index[self._q_valid]] = o[self._q_valid]
        return out


@dataclass
class _Running:
    req: InferenceRequest
    tokens: List[int]                 # prompt + generated so far
    n_prompt: int
    fed: int = 0                      # tokens already forwarded (== kv_len)
    logprobs: List[float] = field(default_factory=list)
    stream: Optional[object] = None   # TokenStream: every sampled token is published as it is produced
    cancelled: bool = False


class DynamicInferenceEngine:
    """Token-granularity continuous batching: requests join and leave between any two forward steps."""

    def __init__(self, model, tokenizer, max_sequence_length: int = 2560, num_blocks: int = 1024, block_size: int = 16,
                 max_tokens_per_step: int = 256, max_requests: int = 64, random_seed: Optional[int] = None):
        self.model, self.tokenizer = model, tokenizer
        self.context = DynamicInferenceContext(max_sequence_length, num_blocks, block_size, max_tokens_per_step,
                                               max_requests)
        self.waiting: "OrderedDict[str, _Running]" = OrderedDict()
        self.running: "OrderedDict[str, _Running]" = OrderedDict()
        self.completed: "OrderedDict[str, InferenceRequest]" = OrderedDict()
        self._ids = itertools.count()
        self.generator = None
        if random_seed is not None:
            self.generator = torch.Generator().manual_seed(random_seed)
        self.num_preemptions = 0
        self.steps = 0
        self.padded_vocab_size = None
        # continuous serving: requests from concurrent clients join the running batch between two steps
        self._inbox: "queue.Queue" = queue.Queue()
        self._done: Dict[str, threading.Event] = {}
        self._loop_thread: Optional[threading.Thread] = None
        self._stop = False

    # -- API -----------------------------------------------------------------------
    def add_request(self, prompt: str, sampling_params: Optional[SamplingParams] = None,
                    prompt_tokens: Optional[List[int]] = None, stream=None) -> str:
        rid = str(next(self._ids))
        toks = list(prompt_tokens) if prompt_tokens is not None else self.tokenizer.tokenize(prompt)
        req = InferenceRequest(rid, prompt, sampling_params or SamplingParams(), toks, time.time())
        self.waiting[rid] = _Running(req, list(toks), len(toks), stream=stream)
        if stream is not None:
            stream.request_id, stream._cancel = rid, self.cancel_request
        return rid

    def cancel_request(self, rid: str) -> None:
        """Drop a request (a streaming client went away): it leaves the batch at the next step boundary."""
        r = self.running.get(rid) or self.waiting.get(rid)
        if r is not None:
            r.cancelled = True

    def generate_stream(self, prompt: str, sampling_params: Optional[SamplingParams] = None,
                        prompt_tokens: Optional[List[int]] = None):
        """Synchronous token streaming: yields ``{"token", "text", "logprob"}`` as soon as a token is sampled, stepping the
        engine itself (stand-alone use) or reading from the serving loop's stream (``serving()``); other requests keep
        sharing the steps."""
        from .async_stream import TokenStream

        st = TokenStream()
        if self.serving() and threading.current_thread() is not self._loop_thread:
            toks = list(prompt_tokens) if prompt_tokens is not None else self.tokenizer.tokenize(prompt)
            self._inbox.put((prompt, sampling_params or SamplingParams(), toks, [], threading.Event(), st))
            yield from st
            return
        self.add_request(prompt, sampling_params, prompt_tokens, stream=st)
        while not st.finished:
            self.step()
            while not st._q.empty():
                item = st._q.get_nowait()
                if isinstance(item, dict):
                    yield item

    def has_unfinished_requests(self) -> bool:
        return bool(self.waiting) or bool(self.running)

    @classmethod
    def from_args(cls, model, tokenizer, args, config) -> "DynamicInferenceEngine":
        """Size the paged cache from ``--inference-dynamic-batching-buffer-size-gb`` (reference
        ``arguments.py:1150-1177``): bytes per token = 2 (K,V) x layers x groups/tp x head_dim x dtype size."""
        bs = args.inference_dynamic_batching_chunk_size
        elt = 2 if (getattr(args, "bf16", False) or getattr(args, "fp16", False)) else 4
        tp = max(1, getattr(args, "tensor_model_parallel_size", 1))
        pp = max(1, getattr(args, "pipeline_model_parallel_size", 1))
        per_tok = 2 * (config.num_layers // pp) * max(1, config.num_query_groups // tp) * config.kv_channels * elt
        blocks = max(2, int(args.inference_dynamic_batching_buffer_size_gb * (1 << 30)) // (per_tok * bs))
        over = args.inference_dynamic_batching_buffer_overflow_factor or 1.0
        max_req = args.inference_dynamic_batching_max_requests_override or max(1, int(blocks * over))
        max_tok = args.inference_dynamic_batching_max_tokens_override or min(16384, int(blocks * bs * over))
        eng = cls(model, tokenizer, args.inference_max_seq_length, blocks, bs, max_tok, max_req, getattr(args, "seed", None))
        eng.padded_vocab_size = getattr(args, "padded_vocab_size", None)
        return eng

    @property
    def controller(self):            # the servers reach the tokenizer through engine.controller.tokenizer
        return self

    # -- continuous serving ------------------------------------------------------------------
    def serving(self) -> bool:
        return self._loop_thread is not None

    def start_background_loop(self):
        """Rank 0 of a server: run ``run_loop`` in a thread; ``generate`` then enqueues instead of driving the engine."""
        self._loop_thread = threading.Thread(target=self.run_loop, daemon=True)
        self._loop_thread.start()

    def stop_loop(self):
        self._stop = True
        if self._loop_thread is not None and self._loop_thread is not threading.current_thread():
            self._loop_thread.join(timeout=30)
        self._loop_thread = None

    def run_loop(self):
        """The serving loop, executed by EVERY model-parallel rank in lockstep: rank 0 drains the inbox and broadcasts
        the newly arrived requests, all ranks admit them in the same order (hence the same request ids) and take one
        engine step.  Requests therefore join and leave the batch at token granularity across HTTP clients."""
        import torch.distributed as dist

        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        rank0 = not multi or dist.get_rank() == 0
        while True:
            new = []
            if rank0:
                try:
                    block = not self.has_unfinished_requests()
                    new.append(self._inbox.get(timeout=0.005) if block else self._inbox.get_nowait())
                    while True:
                        new.append(self._inbox.get_nowait())
                except queue.Empty:
                    pass
            box = [[(p, sp, toks) for p, sp, toks, *_ in new], self._stop]
            streams = [(n[5] if len(n) > 5 else None) for n in new]
            if multi:
                dist.broadcast_object_list(box, src=0)
            if box[1]:
                return
            for i, (prompt, sp, toks) in enumerate(box[0]):
                rid = self.add_request(prompt, sp, toks, stream=streams[i] if (rank0 and i < len(streams)) else None)
                if rank0:
                    new[i][3].append(rid)                      # hand the id back to the submitting thread
                    self._done[rid] = new[i][4]
            if self.has_unfinished_requests():
                for r in self.step():
                    ev = self._done.pop(r.request_id, None)
                    if ev is not None:
                        ev.set()

    def submit(self, prompts: List[str], sampling_params: Optional[SamplingParams], token_lists: List[List[int]],
               timeout: float = 600.0) -> List[InferenceRequest]:
        tickets = []
        for p, toks in zip(prompts, token_lists):
            ids, ev = [], threading.Event()
            self._inbox.put((p, sampling_params or SamplingParams(), toks, ids, ev))
            tickets.append((ids, ev))
        out = []
        for ids, ev in tickets:
            if not ev.wait(timeout):
                raise TimeoutError("generation did not finish in time")
            out.append(self.completed[ids[0]])
        return out

    def generate(self, prompts: List[str], sampling_params: Optional[SamplingParams] = None,
                 add_BOS: bool = False) -> List[InferenceRequest]:
        if self.serving() and threading.current_thread() is not self._loop_thread:
            b = getattr(self.tokenizer, "bos", None) if add_BOS else None
            toks = [([b] if b is not None else []) + self.tokenizer.tokenize(p) for p in prompts]
            return self.submit(prompts, sampling_params, toks)
        bos = getattr(self.tokenizer, "bos", None) if add_BOS else None
        if bos is not None:
            ids = [self.add_request(p, sampling_params, [bos] + self.tokenizer.tokenize(p)) for p in prompts]
            while self.has_unfinished_requests():
                self.step()
            return [self.completed[i] for i in ids]
        ids = [self.add_request(p, sampling_params) for p in prompts]
        while self.has_unfinished_requests():
            self.step()
        return [self.completed[i] for i in ids]

    # -- one step -----------------------------------------------------------------
    def _admit(self):
        ctx = self.context
        while self.waiting and len(self.running) < ctx.max_requests:
            rid, r = next(iter(self.waiting.items()))
            # admit only if the whole (re)prefill plus one decode block fits next to what is already promised
            promised = sum(ctx.blocks_needed(x, len(rr.tokens) - rr.fed + 1) for x, rr in self.running.items())
            if -(-(len(r.tokens) + 1) // ctx.block_size) + promised > ctx.allocator.free:
                break
            self.waiting.pop(rid)
            ctx.add_request(rid)
            r.fed = 0
            r.req.status = Status.ACTIVE_BUT_NOT_GENERATING_TOKENS
            self.running[rid] = r

    def _preempt_one(self):
        """Out of blocks mid-flight: evict the most recently admitted request; it re-prefills later."""
        rid, r = self.running.popitem(last=True)
        self.context.release(rid)
        r.fed = 0
        r.req.status = Status.WAITING_IN_QUEUE
        self.waiting[rid] = r
        self.waiting.move_to_end(rid, last=False)
        self.num_preemptions += 1

    def _plan(self) -> List[tuple]:
        ctx = self.context
        while True:
            budget, work, need = ctx.max_tokens_per_step, [], 0
            # decode tokens first (latency), then prefill chunks with what is left of the budget
            order = sorted(self.running.items(), key=lambda kv: len(kv[1].tokens) - kv[1].fed)
            for rid, r in order:
                n = min(len(r.tokens) - r.fed, budget)
                if n <= 0:
                    continue
                work.append((rid, n))
                need += ctx.blocks_needed(rid, n)
                budget -= n
            if need <= ctx.allocator.free or len(self.running) <= 1:
                return work
            self._preempt_one()

    @torch.no_grad()
    def step(self) -> List[InferenceRequest]:
        self._admit()
        work = self._plan()
        if not work:
            return []
        ctx = self.context
        dev = next(self.model.parameters()).device
        self.model.eval()
        positions = ctx.begin_step(work, dev)
        ids = torch.tensor([t for rid, n in work for t in self.running[rid].tokens[self.running[rid].fed:self.running[rid].fed + n]],
                           dtype=torch.long, device=dev).unsqueeze(0)
        inner = self.model
        while hasattr(inner, "module"):
            inner = inner.module
        out = pipeline_forward(self.model, inner, ids, positions.unsqueeze(0), ctx, (ids.shape[1], 1))
        out = gather_full_vocab(out, self.padded_vocab_size or getattr(inner, "vocab_size", None))
        logits = out[0] if out is not None else None                                          # [T, V] on the last stage
        last = ctx.last_token_index()
        ctx.end_step()
        self.steps += 1
        finished = []
        vocab = getattr(self.tokenizer, "vocab_size", None)
        eod = self.tokenizer.eod
        sampling = [rid for rid, n in work if self.running[rid].fed + n >= len(self.running[rid].tokens)]
        picked = None
        if logits is not None and sampling:                            # last pipeline stage samples ...
            picked = torch.stack([sample_from_logits(logits[last[rid]].unsqueeze(0), self.running[rid].req.inference_parameters,
                                                     vocab, self.generator).view(()) for rid in sampling]).to(dev)
        if sampling:                                                   # ... every model-parallel rank gets the same tokens
            picked = sync_sampled_tokens(picked, len(sampling), dev)
        chosen = dict(zip(sampling, picked.tolist())) if sampling else {}
        for rid, n in work:
            r = self.running[rid]
            r.fed += n
            if r.fed < len(r.tokens):
                continue                                              # still prefilling
            p = r.req.inference_parameters
            new = int(chosen[rid])
            if p.return_log_probs and logits is not None:
                r.logprobs.append(float(torch.log_softmax(logits[last[rid]].float(), -1)[new]))
            r.tokens.append(new)
            r.req.status = Status.ACTIVE_AND_GENERATING_TOKENS
            n_gen = len(r.tokens) - r.n_prompt
            if r.stream is not None and new != eod:
                r.stream.put({"token": new, "text": self.tokenizer.detokenize([new]),
                              "logprob": r.logprobs[-1] if (p.return_log_probs and r.logprobs) else None})
            if new == eod or n_gen >= p.num_tokens_to_generate or len(r.tokens) >= ctx.max_sequence_length or r.cancelled:
                gen = r.tokens[r.n_prompt:]
                r.req.generated_tokens = torch.tensor(gen, dtype=torch.long)
                r.req.generated_length = len(gen)
                r.req.generated_log_probs = r.logprobs if p.return_log_probs else None
                r.req.generated_text = self.tokenizer.detokenize(gen[:-1] if gen and gen[-1] == eod else gen)
                r.req.status = Status.COMPLETED
                ctx.release(rid)
                self.running.pop(rid)
                self.completed[rid] = r.req
                finished.append(r.req)
                if r.stream is not None:
                    r.stream.finish(r.req)
        return finished


class ContextOverflowError(Exception):
    """The dynamic context cannot take a request now (reference ``contexts/dynamic_context.py``)."""


class RequestOverflowError(ContextOverflowError):
    """More concurrent requests than ``max_requests``."""


class TokenOverflowError(ContextOverflowError):
    """More tokens in one step than ``max_tokens_per_step``."""


class ChunkOverflowError(ContextOverflowError):
    """No free KV block (the engine pauses or evicts a request instead of raising)."""
