"""``megatron.core.optimizer.clip_grads``: gradient norm / clipping / zero counting over lists of tensors, on the multi-tensor
kernels (``ops.l2norm_sq`` / ``ops.scale_``: no host synchronisation before the final scalar)."""
from __future__ import annotations

import math
from typing import List, Optional, Union

import torch
import torch.distributed as dist

from .. import ops


def _real(group) -> bool:
    return group is not None and dist.is_available() and dist.is_initialized() and not type(group).__name__.startswith("_Single")


def get_grad_norm_fp32(grads_for_norm: Union[List[torch.Tensor], torch.Tensor], norm_type: Union[int, float] = 2,
                       grad_stats_parallel_group=None, model_parallel_group=None) -> float:
    """Norm over ``grads_for_norm`` (this rank's non-duplicated gradients), reduced over the group that owns disjoint shards."""
    group = grad_stats_parallel_group if grad_stats_parallel_group is not None else model_parallel_group
    grads = [grads_for_norm] if isinstance(grads_for_norm, torch.Tensor) else [g for g in grads_for_norm if g is not None]
    dev = grads[0].device if grads else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    if norm_type == math.inf:
        total = torch.stack([g.detach().abs().max().float() for g in grads]).max() if grads else torch.zeros((), device=dev)
        total = total.reshape(1)
        if _real(group):
            dist.all_reduce(total, op=dist.ReduceOp.MAX, group=group)
        return float(total.item())
    if norm_type == 2:
        total = ops.l2norm_sq([g.detach() for g in grads]).reshape(1) if grads else torch.zeros(1, device=dev)
    else:
        total = sum((g.detach().float().abs() ** norm_type).sum() for g in grads).reshape(1) if grads else torch.zeros(1, device=dev)
    total = total.float()
    if _real(group):
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return float(total.item()) ** (1.0 / norm_type)


def clip_grad_by_total_norm_fp32(parameters: Union[List[torch.Tensor], torch.Tensor], max_norm: Union[int, float],
                                 total_norm: float, use_decoupled_grad: bool = False) -> None:
    params = [parameters] if isinstance(parameters, torch.Tensor) else list(parameters)
    attr = "decoupled_grad" if use_decoupled_grad else "grad"
    clip_coeff = max_norm / (total_norm + 1.0e-6)
    if clip_coeff >= 1.0:
        return
    for p in params:
        g = getattr(p, attr, None)
        if g is not None:
            ops.scale_(g.detach(), clip_coeff)


def count_zeros_fp32(parameters: Union[List[torch.Tensor], torch.Tensor], grad_stats_parallel_group=None,
                     use_decoupled_grad: bool = False, model_parallel_group=None) -> float:
    group = grad_stats_parallel_group if grad_stats_parallel_group is not None else model_parallel_group
    params = [parameters] if isinstance(parameters, torch.Tensor) else list(parameters)
    attr = "decoupled_grad" if use_decoupled_grad else "grad"
    dev = params[0].device if params else torch.device("cpu")
    total = torch.zeros(1, dtype=torch.float32, device=dev)
    for p in params:
        g = getattr(p, attr, None)
        shared = getattr(p, "shared", False)
        if g is not None and not shared and not getattr(p, "_tp_duplicate", False):
            total += g.numel() - torch.count_nonzero(g.detach())
    if _real(group):
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return float(total.item())
