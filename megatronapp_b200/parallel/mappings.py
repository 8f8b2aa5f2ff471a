"""Autograd-aware collectives between tensor-parallel / sequence-parallel regions.

API parity with reference ``megatron/core/tensor_parallel/mappings.py:354-732`` (function
names kept); each collective is wrapped in a MegaScan scope tagged with the participating
ranks (reference ``mappings.py:32-49``) through one helper instead of duplicated traced /
untraced bodies.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import state as ps
from ..scan.tracer import scope as _scope


def _tp():
    return ps.get_tensor_model_parallel_group(), ps.get_tensor_model_parallel_world_size()


def _group_ranks():
    try:
        return ps.get_group_ranks("tp")
    except Exception:
        return None


# ---- raw collectives ----------------------------------------------------------

def _reduce(x, group=None):
    group = group or ps.get_tensor_model_parallel_group()
    if ps.get_tensor_model_parallel_world_size() == 1:
        return x
    with _scope("_reduce", group=_group_ranks, expect="_reduce", data=x.numel() * x.element_size()):
        dist.all_reduce(x, group=group)
    return x


def _split_along_last_dim(x):
    g, ws = _tp()
    if ws == 1:
        return x
    rank = ps.get_tensor_model_parallel_rank()
    return x.chunk(ws, dim=-1)[rank].contiguous()


def _split_along_first_dim(x):
    g, ws = _tp()
    if ws == 1:
        return x
    rank = ps.get_tensor_model_parallel_rank()
    n = x.shape[0] // ws
    return x[rank * n:(rank + 1) * n].contiguous()


def _gather_along_last_dim(x):
    g, ws = _tp()
    if ws == 1:
        return x
    with _scope("_gather_along_last_dim", group=_group_ranks, expect="_gather_along_last_dim",
                data=x.numel() * x.element_size() * ws):
        parts = [torch.empty_like(x) for _ in range(ws)]
        dist.all_gather(parts, x.contiguous(), group=g)
    return torch.cat(parts, dim=-1).contiguous()


def _gather_along_first_dim(x, group=None, output_split_sizes=None):
    g = group or ps.get_tensor_model_parallel_group()
    ws = 1 if g is None else dist.get_world_size(g) if dist.is_initialized() and not isinstance(g, ps._SingleGroup) else 1
    if ws == 1:
        return x
    with _scope("_gather_along_first_dim", group=_group_ranks, expect="_gather_along_first_dim",
                data=x.numel() * x.element_size() * ws):
        if output_split_sizes is None:
            out = torch.empty((x.shape[0] * ws,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            dist.all_gather_into_tensor(out, x.contiguous(), group=g)
        else:
            out = torch.empty((sum(output_split_sizes),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            dist.all_gather(list(torch.split(out, output_split_sizes, dim=0)), x.contiguous(), group=g)
    return out


def _reduce_scatter_along_first_dim(x, group=None, input_split_sizes=None):
    g = group or ps.get_tensor_model_parallel_group()
    ws = 1 if g is None else dist.get_world_size(g) if dist.is_initialized() and not isinstance(g, ps._SingleGroup) else 1
    if ws == 1:
        return x
    with _scope("_reduce_scatter_along_first_dim", group=_group_ranks,
                expect="_reduce_scatter_along_first_dim", data=x.numel() * x.element_size()):
        if input_split_sizes is None:
            out = torch.empty((x.shape[0] // ws,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            dist.reduce_scatter_tensor(out, x.contiguous(), group=g)
        else:
            rank = dist.get_rank(g)
            out = torch.empty((input_split_sizes[rank],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            dist.reduce_scatter(out, list(torch.split(x.contiguous(), input_split_sizes, dim=0)), group=g)
    return out


def _reduce_scatter_along_last_dim(x):
    g, ws = _tp()
    if ws == 1:
        return x
    with _scope("_reduce_scatter_along_last_dim", group=_group_ranks,
                expect="_reduce_scatter_along_last_dim", data=x.numel() * x.element_size()):
        parts = [p.contiguous() for p in x.chunk(ws, dim=-1)]
        out = torch.empty_like(parts[0])
        dist.reduce_scatter(out, parts, group=g)
    return out


# ---- autograd wrappers ----------------------------------------------------------

class _CopyToModelParallelRegion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x

    @staticmethod
    def backward(ctx, g):
        return _reduce(g.contiguous())


class _ReduceFromModelParallelRegion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _reduce(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return g


class _ScatterToModelParallelRegion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _split_along_last_dim(x)

    @staticmethod
    def backward(ctx, g):
        return _gather_along_last_dim(g)


class _GatherFromModelParallelRegion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _gather_along_last_dim(x)

    @staticmethod
    def backward(ctx, g):
        return _split_along_last_dim(g)


class _ScatterToSequenceParallelRegion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _split_along_first_dim(x)

    @staticmethod
    def backward(ctx, g):
        return _gather_along_first_dim(g)


class _GatherFromSequenceParallelRegion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, tensor_parallel_output_grad, group, output_split_sizes):
        ctx.tpo, ctx.group, ctx.splits = tensor_parallel_output_grad, group, output_split_sizes
        return _gather_along_first_dim(x, group, output_split_sizes)

    @staticmethod
    def backward(ctx, g):
        if ctx.tpo:
            return _reduce_scatter_along_first_dim(g, ctx.group, ctx.splits), None, None, None
        return _split_along_first_dim(g), None, None, None


class _ReduceScatterToSequenceParallelRegion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group, input_split_sizes):
        ctx.group, ctx.splits = group, input_split_sizes
        return _reduce_scatter_along_first_dim(x, group, input_split_sizes)

    @staticmethod
    def backward(ctx, g):
        return _gather_along_first_dim(g, ctx.group, ctx.splits), None, None


class _AllGatherFromTensorParallelRegion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _gather_along_last_dim(x)

    @staticmethod
    def backward(ctx, g):
        return _reduce_scatter_along_last_dim(g)


class _ReduceScatterToTensorParallelRegion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _reduce_scatter_along_last_dim(x)

    @staticmethod
    def backward(ctx, g):
        return _gather_along_last_dim(g)


class _AllToAll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, group, x, output_split_sizes, input_split_sizes):
        ctx.group, ctx.out_s, ctx.in_s = group, output_split_sizes, input_split_sizes
        ws = dist.get_world_size(group) if (group is not None and dist.is_initialized()
                                            and not isinstance(group, ps._SingleGroup)) else 1
        if ws == 1:
            return x
        x = x.contiguous()
        if output_split_sizes is None:
            out = torch.empty_like(x)
        else:
            out = torch.empty((sum(output_split_sizes),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        with _scope("all_to_all", data=x.numel() * x.element_size()):
            dist.all_to_all_single(out, x, output_split_sizes=output_split_sizes,
                                   input_split_sizes=input_split_sizes, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        return None, _AllToAll.apply(ctx.group, g, ctx.in_s, ctx.out_s), None, None


def copy_to_tensor_model_parallel_region(x):
    return _CopyToModelParallelRegion.apply(x)


def reduce_from_tensor_model_parallel_region(x):
    return _ReduceFromModelParallelRegion.apply(x)


def scatter_to_tensor_model_parallel_region(x):
    return _ScatterToModelParallelRegion.apply(x)


def gather_from_tensor_model_parallel_region(x):
    return _GatherFromModelParallelRegion.apply(x)


def scatter_to_sequence_parallel_region(x):
    return _ScatterToSequenceParallelRegion.apply(x)


def gather_from_sequence_parallel_region(x, tensor_parallel_output_grad=True, group=None,
                                         output_split_sizes=None):
    return _GatherFromSequenceParallelRegion.apply(x, tensor_parallel_output_grad, group, output_split_sizes)


def reduce_scatter_to_sequence_parallel_region(x, group=None, input_split_sizes=None):
    return _ReduceScatterToSequenceParallelRegion.apply(x, group, input_split_sizes)


def all_gather_last_dim_from_tensor_parallel_region(x):
    return _AllGatherFromTensorParallelRegion.apply(x)


def reduce_scatter_last_dim_to_tensor_parallel_region(x):
    return _ReduceScatterToTensorParallelRegion.apply(x)


def all_to_all(group, x, output_split_sizes=None, input_split_sizes=None):
    return _AllToAll.apply(group, x, output_split_sizes, input_split_sizes)


def all_to_all_sp2hp(x):
    """[s/tp, b, h] -> [s, b, h/tp]  (Ulysses-style; reference mappings.py:685)."""
    g, ws = _tp()
    if ws == 1:
        return x
    s_local, b, h = x.shape
    x = x.reshape(s_local * b, h)
    parts = torch.cat([p.contiguous() for p in x.chunk(ws, dim=-1)], dim=0)
    out = all_to_all(g, parts)
    return out.reshape(s_local * ws, b, h // ws)


def all_to_all_hp2sp(x):
    """[s, b, h/tp] -> [s/tp, b, h]  (reference mappings.py:710)."""
    g, ws = _tp()
    if ws == 1:
        return x
    s, b, hp = x.shape
    x = x.reshape(s * b, hp)
    out = all_to_all(g, x)
    parts = out.chunk(ws, dim=0)
    return torch.cat(parts, dim=-1).reshape(s // ws, b, hp * ws)
