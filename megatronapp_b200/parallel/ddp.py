"""Bucketed data-parallel gradient reduction (DDP and the ZeRO-1 reduce-scatter half).

Parity: reference ``core/distributed/distributed_data_parallel.py`` (520) +
``param_and_grad_buffer.py`` (885) + ``finalize_model_grads.py`` (331): contiguous per-dtype
param / grad buffers, ``param.main_grad`` views, reverse-order bucketing (default
``max(40M, 1M*dp)`` elements, padded to ``lcm(dp, 128)`` for the sharded optimizer), grad-ready
hooks that launch a bucket's reduction as soon as its last gradient of the last microbatch
exists, ``no_sync()``, ``1/dp`` scaling (or ``AVG``), expert params reduced over the expert-DP
group, and the finalisation order (DP sync -> LN/SP grads AR over TP -> embedding AR first<->last
stage -> per-token loss scaling).

B200-first differences: (1) params are *views into one contiguous buffer*, so the optimizer is a
single fused Adam launch per buffer range; (2) ready-notification for GEMM weights comes
straight from the wgrad kernel's host wrapper (``weight.grad_ready_hook``), no dummy gradients;
(3) with a DP symmetric arena the bucket reduce-scatter is ONE sm_100a kernel fused with the
scale and the wire cast over NVLink peer stores (``bucket_reduce_scatter``) instead of
``torch.mul`` + NCCL; (4) the reference's undefined-``cm`` bug has no analogue.
"""
from __future__ import annotations

import os
import math
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import state as ps
from .. import ops
from ..scan.tracer import scope as _scope


@dataclass
class DistributedDataParallelConfig:
    grad_reduce_in_fp32: bool = False
    overlap_grad_reduce: bool = False
    overlap_param_gather: bool = False
    align_param_gather: bool = False
    use_distributed_optimizer: bool = False
    num_distributed_optimizer_instances: int = 1
    check_for_nan_in_grad: bool = False
    check_for_large_grads: bool = False
    bucket_size: Optional[int] = None
    average_in_collective: bool = False
    fp8_param_gather: bool = False
    use_custom_fsdp: bool = False
    data_parallel_sharding_strategy: str = "no_shard"
    gradient_reduce_div_fusion: bool = True
    keep_fp8_transpose_cache_when_using_custom_fsdp: bool = False
    # B200: fused scale/cast + reduce-scatter kernel over the DP symmetric arena
    fused_grad_reduce_scatter: bool = False
    fused_rs_ctas: int = int(os.environ.get("B200_DP_RS_CTAS", "48"))   # SM carve-out of the fused bucket reduce-scatter
    grad_wire_dtype: str = "fp32"  # fp32 | bf16


def _is_real_group(g) -> bool:
    return g is not None and not isinstance(g, ps._SingleGroup) and dist.is_initialized() and dist.get_world_size(g) > 1


class _Bucket:
    def __init__(self, params: List[nn.Parameter], start: int, end: int, bucket_id: int):
        self.params, self.start, self.end, self.bucket_id = params, start, end, bucket_id
        self.params_set = set(params)
        self.ready = set()
        self.handle = None
        self.dispatched = False


class ParamAndGradBuffer:
    """One contiguous (param, grad) pair of buffers for params of a given (param dtype, grad dtype)."""

    def __init__(self, ddp_config, param_dtype, grad_dtype, params: List[nn.Parameter], group, group_size: int,
                 bucket_size: Optional[int], gradient_scaling_factor: float, device, arena_kind: Optional[str] = None):
        self.ddp_config, self.param_dtype, self.grad_dtype = ddp_config, param_dtype, grad_dtype
        self.group, self.group_size = group, group_size
        self.gradient_scaling_factor = gradient_scaling_factor
        self.params = list(params)
        self.is_last_microbatch = True
        self.arena_kind = arena_kind
        pad_to = 1
        if ddp_config.use_distributed_optimizer:
            pad_to = math.lcm(group_size, 128) * (4 if ddp_config.fused_grad_reduce_scatter else 1)

        def pad(n, m):
            return (n + m - 1) // m * m

        # reverse order: buckets fill in (approximately) backward-pass order
        self.param_index: Dict[nn.Parameter, tuple] = {}
        self.buckets: List[_Bucket] = []
        off = 0
        bstart, bparams = 0, []
        for p in reversed(self.params):
            n = p.numel()
            own_bucket = getattr(p, "shared_embedding", False) and ddp_config.use_distributed_optimizer
            if own_bucket and bparams:
                # the tied embedding must sit ALONE in its bucket on every pipeline stage (flush what
                # came before it as well as closing after it, reference param_and_grad_buffer.py:582-590):
                # the first<->last stage embedding all-reduce runs on reduce-scattered shards, so both
                # stages must own identical element ranges of the weight
                off = pad(off, pad_to)
                self.buckets.append(_Bucket(bparams, bstart, off, len(self.buckets)))
                bstart, bparams = off, []
            start = pad(off, 64)  # 128B/256B-aligned views: TMA + vectorised kernels need 16B alignment
            end = start + n
            self.param_index[p] = (start, end, len(self.buckets))
            bparams.append(p)
            off = end
            if (bucket_size is not None and (off - bstart) >= bucket_size) or own_bucket:
                off = pad(off, pad_to)
                self.buckets.append(_Bucket(bparams, bstart, off, len(self.buckets)))
                bstart, bparams = off, []
        if bparams:
            off = pad(off, pad_to)
            self.buckets.append(_Bucket(bparams, bstart, off, len(self.buckets)))
        self.numel = off
        self.grad_data = torch.zeros(self.numel, dtype=grad_dtype, device=device)
        self.param_data = torch.zeros(self.numel, dtype=param_dtype, device=device)
        for p in self.params:
            s, e, _ = self.param_index[p]
            view = self.param_data[s:e].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.main_grad = self.grad_data[s:e].view(p.shape)
            p.grad_added_to_main_grad = False
        self.param_to_bucket = {p: self.buckets[self.param_index[p][2]] for p in self.params}
        self._rs_epoch = 0
        # params whose first wgrad of a step is a plain store (epilogue "=") need no zeroing
        skip = sorted((self.param_index[p][0], self.param_index[p][1]) for p in self.params
                      if getattr(p, "b200_wgrad_store_first", False))
        self._zero_ranges = None
        if skip:
            rng, cur = [], 0
            for s0, e0 in skip:
                if s0 > cur:
                    rng.append((cur, s0))
                cur = e0
            if cur < self.numel:
                rng.append((cur, self.numel))
            self._zero_ranges = rng

    # ---- gradient sync ---------------------------------------------------------
    def reset(self):
        for b in self.buckets:
            b.ready.clear()
            b.handle = None
            b.dispatched = False

    def zero(self):
        if self._zero_ranges is None:
            self.grad_data.zero_()
        else:
            for a, b in self._zero_ranges:
                self.grad_data[a:b].zero_()
        for p in self.params:
            p.grad_added_to_main_grad = False
        self.reset()

    def shard_range(self, bucket: _Bucket, rank: Optional[int] = None):
        r = dist.get_rank(self.group) if (rank is None and _is_real_group(self.group)) else (rank or 0)
        n = (bucket.end - bucket.start) // max(1, self.group_size)
        return bucket.start + r * n, bucket.start + (r + 1) * n

    def start_bucket_sync(self, bucket: _Bucket, async_op: bool):
        if bucket.dispatched:
            return
        bucket.dispatched = True
        g = self.grad_data[bucket.start:bucket.end]
        if self.ddp_config.check_for_nan_in_grad:
            if not torch.isfinite(g.float().sum()):
                raise RuntimeError(f"NaN/Inf in gradients of bucket {bucket.bucket_id} before reduction")
        if self.ddp_config.check_for_large_grads:
            # reference param_and_grad_buffer.py:180-190: a bucket norm 10x the largest seen so far is reported / re-run
            from functools import partial

            from ..training.rerun_state_machine import get_rerun_state_machine

            rsm = get_rerun_state_machine()
            rsm.validate_result(g.float().norm(), partial(rsm.is_unexpectedly_large, threshold=10, context="grads"),
                                "found unexpected large grads in bucket before data-parallel communication collective",
                                tolerance=0.001, fatal=False)
        if not _is_real_group(self.group):
            if self.gradient_scaling_factor != 1.0:
                ops.scale_(g, self.gradient_scaling_factor)
            return
        cfg = self.ddp_config
        with _scope("grad-reduce-scatter" if cfg.use_distributed_optimizer else "allreduce",
                    data=g.numel() * g.element_size()):
            if cfg.use_distributed_optimizer and cfg.fused_grad_reduce_scatter and self._fused_rs(bucket, g):
                return
            op = dist.ReduceOp.SUM
            if cfg.average_in_collective:
                op = dist.ReduceOp.AVG
            elif self.gradient_scaling_factor != 1.0:
                ops.scale_(g, self.gradient_scaling_factor)
            if cfg.use_distributed_optimizer:
                s, e = self.shard_range(bucket)
                bucket.handle = dist.reduce_scatter_tensor(self.grad_data[s:e], g, op=op, group=self.group,
                                                           async_op=async_op)
            else:
                bucket.handle = dist.all_reduce(g, op=op, group=self.group, async_op=async_op)

    def _fused_rs(self, bucket: _Bucket, g: torch.Tensor) -> bool:
        from ..comm.arena import get_arena

        arena = get_arena(self.arena_kind or "dp")
        if arena is None or g.dtype != torch.float32 or not g.is_cuda:
            return False
        n = g.numel()
        wire = 2 if self.ddp_config.grad_wire_dtype == "fp32" else 0
        wire_bytes = 4 if wire == 2 else 2
        slot = self._rs_epoch & 1
        stage = arena.region(f"dp_rs_stage{slot}", self.max_bucket_numel() * wire_bytes)
        flags = arena.region(f"dp_rs_flags{slot}", 128 * 8 * 4)
        epoch = arena.next_epoch(flags)
        self._rs_epoch += 1
        s, e = self.shard_range(bucket, arena.rank)
        scale = self.gradient_scaling_factor if not self.ddp_config.average_in_collective else 1.0 / self.group_size
        # side stream: the 32-64 CTA reduce-scatter runs next to the backward GEMMs instead of in front of them; the
        # optimizer joins it in finish_grad_sync
        cur = torch.cuda.current_stream()
        if self.ddp_config.overlap_grad_reduce:
            if getattr(self, "_rs_stream", None) is None:
                self._rs_stream = torch.cuda.Stream()
            self._rs_stream.wait_stream(cur)
            with torch.cuda.stream(self._rs_stream):
                ops.ext().bucket_reduce_scatter(g, self.grad_data[s:e], arena.ptrs(stage), arena.ptrs(flags), arena.rank,
                                                arena.size, scale, epoch, wire, self.ddp_config.fused_rs_ctas)
            self._rs_pending = True
        else:
            ops.ext().bucket_reduce_scatter(g, self.grad_data[s:e], arena.ptrs(stage), arena.ptrs(flags), arena.rank,
                                            arena.size, scale, epoch, wire, self.ddp_config.fused_rs_ctas)
        return True

    def max_bucket_numel(self) -> int:
        return max(b.end - b.start for b in self.buckets)

    def register_grad_ready(self, param: nn.Parameter):
        bucket = self.param_to_bucket[param]
        bucket.ready.add(param)
        if len(bucket.ready) == len(bucket.params_set):
            self.start_bucket_sync(bucket, async_op=True)

    def start_grad_sync(self, async_op: bool = False):
        for b in self.buckets:
            self.start_bucket_sync(b, async_op)

    def finish_grad_sync(self):
        for b in self.buckets:
            if not b.dispatched:
                self.start_bucket_sync(b, async_op=self.ddp_config.overlap_grad_reduce)
        for b in self.buckets:
            if b.handle is not None:
                b.handle.wait()
                b.handle = None
        if getattr(self, "_rs_pending", False):
            torch.cuda.current_stream().wait_stream(self._rs_stream)
            self._rs_pending = False


class DistributedDataParallel(nn.Module):
    def __init__(self, config, ddp_config: DistributedDataParallelConfig, module: nn.Module,
                 disable_bucketing: bool = False):
        super().__init__()
        self.config, self.ddp_config, self.module = config, ddp_config, module
        dp = ps.get_data_parallel_world_size(with_context_parallel=True)
        bucket_size = ddp_config.bucket_size
        if bucket_size is None and ddp_config.overlap_grad_reduce:
            bucket_size = max(40_000_000, 1_000_000 * dp)
        if not ddp_config.overlap_grad_reduce and not ddp_config.use_distributed_optimizer:
            bucket_size = None
        if disable_bucketing or ps.get_pipeline_model_parallel_rank() > 0:
            # later pipeline stages finish their backward first; a single bucket is enough there
            bucket_size = bucket_size if not disable_bucketing else None
        self.bucket_size = bucket_size
        dense, expert = [], []
        for name, p in module.named_parameters():
            if not p.requires_grad:
                continue
            p.grad_added_to_main_grad = False
            (dense if getattr(p, "allreduce", True) else expert).append(p)
        device = next(module.parameters()).device
        if config.calculate_per_token_loss:
            scale = 1.0
        elif ddp_config.average_in_collective:
            scale = 1.0
        else:
            scale = 1.0 / dp
        self.buffers: List[ParamAndGradBuffer] = []
        self.expert_parallel_buffers: List[ParamAndGradBuffer] = []
        st = ps.get_state()
        dp_group = ps.get_data_parallel_group(with_context_parallel=True)
        self._make_buffers(dense, dp_group, dp, scale, device, self.buffers, "dp")
        if expert:
            edp = st.world("edp")
            escale = 1.0 if config.calculate_per_token_loss else (1.0 / edp if not ddp_config.average_in_collective
                                                                  else 1.0)
            self._make_buffers(expert, st.group("edp"), edp, escale, device, self.expert_parallel_buffers, "edp")
        self.param_to_buffer = {}
        for buf in self.buffers + self.expert_parallel_buffers:
            for p in buf.params:
                self.param_to_buffer[p] = buf
        self._hooks = []
        for p, buf in self.param_to_buffer.items():
            self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(buf)))
            p.grad_ready_hook = self._make_fused_ready(p, buf)
        self.is_last_microbatch = True

    def _make_buffers(self, params, group, group_size, scale, device, out, kind):
        by_dtype: Dict[tuple, List[nn.Parameter]] = {}
        for p in params:
            gdt = torch.float32 if self.ddp_config.grad_reduce_in_fp32 else p.dtype
            by_dtype.setdefault((p.dtype, gdt), []).append(p)
        for (pdt, gdt), ps_ in by_dtype.items():
            out.append(ParamAndGradBuffer(self.ddp_config, pdt, gdt, ps_, group, group_size, self.bucket_size,
                                          scale, device, arena_kind=kind))

    def _make_hook(self, buf: ParamAndGradBuffer):
        def hook(param):
            if param.grad is not None:
                # the fused wgrad path returns None (never a dummy), so whatever arrives here is a
                # real contribution (e.g. the embedding lookup of a tied weight) and is always added
                if param.main_grad.dtype == torch.float32 and param.grad.is_cuda and param.grad.is_contiguous():
                    ops.ext().add_into_f32(param.main_grad, param.grad)
                else:
                    param.main_grad.add_(param.grad.to(param.main_grad.dtype))
                param.grad_added_to_main_grad = True
                param.grad = None
            if self.ddp_config.overlap_grad_reduce and self.is_last_microbatch:
                buf.register_grad_ready(param)
        return hook

    def _make_fused_ready(self, param, buf: ParamAndGradBuffer):
        def ready():
            if self.ddp_config.overlap_grad_reduce and self.is_last_microbatch:
                buf.register_grad_ready(param)
        return ready

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)

    @contextmanager
    def no_sync(self):
        self.is_last_microbatch = False
        try:
            yield
        finally:
            self.is_last_microbatch = True

    def start_grad_sync(self, *unused):
        for buf in self.buffers + self.expert_parallel_buffers:
            buf.start_grad_sync(async_op=self.ddp_config.overlap_grad_reduce)

    def finish_grad_sync(self):
        for buf in self.buffers + self.expert_parallel_buffers:
            buf.finish_grad_sync()

    def scale_gradients(self, scaling_factor: float):
        for buf in self.buffers + self.expert_parallel_buffers:
            ops.scale_(buf.grad_data, scaling_factor)

    def zero_grad_buffer(self):
        for buf in self.buffers + self.expert_parallel_buffers:
            buf.zero()

    def broadcast_params(self):
        for buf in self.buffers:
            if _is_real_group(buf.group):
                dist.broadcast(buf.param_data, src=dist.get_global_rank(buf.group, 0), group=buf.group)
        for buf in self.expert_parallel_buffers:
            if _is_real_group(buf.group):
                dist.broadcast(buf.param_data, src=dist.get_global_rank(buf.group, 0), group=buf.group)

    def set_input_tensor(self, t):
        return self.module.set_input_tensor(t)

    def state_dict(self, *a, **kw):
        return self.module.state_dict(*a, **kw)

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        return self.module.state_dict_for_save_checkpoint(prefix=prefix, keep_vars=keep_vars)

    def sharded_state_dict(self, prefix="", *a, **kw):
        return self.module.sharded_state_dict(prefix, *a, **kw)

    def load_state_dict(self, sd, strict=True):
        return self.module.load_state_dict(sd, strict=strict)


# ---------------------------------------------------------------------------
# finalize_model_grads (reference finalize_model_grads.py:257-331)
# ---------------------------------------------------------------------------

def _unwrap(m):
    while hasattr(m, "module"):
        m = m.module
    return m


def _allreduce_layernorm_grads(model: List[nn.Module], config):
    """Grads of params replicated across TP but fed sequence-sharded data (LN weights under SP,
    q/k layernorm) are partial sums: all-reduce them over the TP group, coalesced."""
    tp = ps.get_tensor_model_parallel_world_size()
    if tp == 1 or not (config.sequence_parallel or config.qk_layernorm):
        return
    grads = []
    for chunk in model:
        for name, p in _unwrap(chunk).named_parameters():
            if p.requires_grad and (getattr(p, "sequence_parallel", False) or "q_layernorm" in name
                                    or "k_layernorm" in name):
                grads.append(p.main_grad if hasattr(p, "main_grad") else p.grad)
    grads = [g for g in grads if g is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    dist.all_reduce(flat, group=ps.get_tensor_model_parallel_group())
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def _allreduce_word_embedding_grads(model: List[nn.Module], config):
    if not (ps.is_rank_in_embedding_group(ignore_virtual=True) and ps.get_pipeline_model_parallel_world_size() > 1):
        return
    g = ps.get_embedding_group()
    if not _is_real_group(g):
        return
    # first stage, last stage, and (encoder-decoder split) the first decoder stage each hold a copy
    m = _unwrap(model[-1]) if ps.is_pipeline_last_stage(ignore_virtual=True) else _unwrap(model[0])
    if getattr(m, "share_embeddings_and_output_weights", False):
        w = m.shared_embedding_or_output_weight()
        grad = w.main_grad if hasattr(w, "main_grad") else w.grad
        dist.all_reduce(grad, group=g)


def _allreduce_position_embedding_grads(model: List[nn.Module], config):
    """Encoder and decoder halves of a pipeline-split T5 share the learned position table (reference
    ``finalize_model_grads._allreduce_position_embedding_grads``)."""
    if ps.get_pipeline_model_parallel_split_rank() is None or not ps.is_rank_in_position_embedding_group():
        return
    g = ps.get_position_embedding_group()
    if not _is_real_group(g):
        return
    emb = getattr(_unwrap(model[0]), "embedding", None)
    pos = getattr(emb, "position_embeddings", None) if emb is not None else None
    if pos is not None:
        w = pos.weight
        dist.all_reduce(w.main_grad if hasattr(w, "main_grad") else w.grad, group=g)


def finalize_model_grads(model: List[nn.Module], num_tokens: Optional[torch.Tensor] = None):
    config = _unwrap(model[0]).config
    with _scope("allreduce"):
        for chunk in model:
            if hasattr(chunk, "finish_grad_sync"):
                chunk.finish_grad_sync()
    _allreduce_layernorm_grads(model, config)
    _allreduce_word_embedding_grads(model, config)
    _allreduce_position_embedding_grads(model, config)
    if num_tokens is not None:
        # per-token loss: broadcast the token count from the last stage, sum over DP, rescale
        last = ps.get_pipeline_model_parallel_last_rank() if ps.model_parallel_is_initialized() else 0
        if _is_real_group(ps.get_pipeline_model_parallel_group()):
            dist.broadcast(num_tokens, src=last, group=ps.get_pipeline_model_parallel_group())
        dpg = ps.get_data_parallel_group(with_context_parallel=True)
        if _is_real_group(dpg):
            dist.all_reduce(num_tokens, group=dpg)
        for chunk in model:
            if num_tokens > 0 and hasattr(chunk, "scale_gradients"):
                chunk.scale_gradients(1.0 / num_tokens.item())
