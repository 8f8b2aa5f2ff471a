"""MoE helpers: permutation, load-balancing losses, capacity, metric tracking.

Parity: reference ``transformer/moe/moe_utils.py`` (``permute`` / ``unpermute`` /
``sort_chunks_by_idxs``, ``switch_load_balancing_loss_func``, ``z_loss_func``,
``sequence_load_balancing_loss_func``, ``get_capacity``, ``MoEAuxLossAutoScaler``,
``track_moe_metrics``, ``topk_softmax_with_capacity`` incl. group-limited routing).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

_MOE_LAYER_WISE_LOGGING_TRACKER: Dict[str, dict] = {}


def switch_load_balancing_loss_func(probs, tokens_per_expert, topk, moe_aux_loss_coeff, sequence_partition_group=None):
    num_sub_sequence = 1
    if sequence_partition_group is not None and dist.is_initialized():
        num_sub_sequence = dist.get_world_size(sequence_partition_group)
        dist.all_reduce(tokens_per_expert, group=sequence_partition_group)
    num_tokens = probs.shape[0] * num_sub_sequence
    num_experts = probs.shape[1]
    aggregated = probs.sum(dim=0)
    return torch.sum(aggregated * tokens_per_expert) * (num_experts * moe_aux_loss_coeff / (num_tokens * num_tokens * topk))


def sequence_load_balancing_loss_func(probs, routing_map, batch_size, seq_length, topk, moe_aux_loss_coeff,
                                      sequence_partition_group=None):
    num_sub = 1
    num_experts = probs.shape[1]
    p = probs.view(seq_length, batch_size, -1)
    r = routing_map.view(seq_length, batch_size, -1)
    if sequence_partition_group is not None and dist.is_initialized():
        num_sub = dist.get_world_size(sequence_partition_group)
        seq_length *= num_sub
    cost = r.sum(dim=0).float()
    if num_sub > 1:
        dist.all_reduce(cost, group=sequence_partition_group)
    cost = cost / (seq_length * topk / num_experts)
    return (p.mean(dim=0) * cost).sum(dim=1).mean() * moe_aux_loss_coeff


def z_loss_func(logits, z_loss_coeff):
    return torch.mean(torch.square(torch.logsumexp(logits, dim=-1))) * z_loss_coeff


def get_capacity(num_tokens: int, num_experts: int, capacity_factor: float, min_capacity=None):
    cap = math.ceil((num_tokens / num_experts) * capacity_factor)
    if min_capacity is not None and cap < min_capacity:
        cap = min_capacity
    return cap


class MoEAuxLossAutoScaler(torch.autograd.Function):
    """Attach an auxiliary loss to the activations so its gradient flows with the main loss."""

    main_loss_backward_scale: torch.Tensor = torch.tensor(1.0)

    @staticmethod
    def forward(ctx, output, aux_loss):
        ctx.save_for_backward(aux_loss)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        (aux_loss,) = ctx.saved_tensors
        scale = MoEAuxLossAutoScaler.main_loss_backward_scale
        return grad_output, torch.ones_like(aux_loss) * scale.to(aux_loss.device).reshape(-1)[0]

    @staticmethod
    def set_loss_scale(scale):
        MoEAuxLossAutoScaler.main_loss_backward_scale = scale


def permute(tokens, routing_map, num_out_tokens: Optional[int] = None, drop_and_pad: bool = False):
    """tokens [n, h], routing_map [n, E] bool -> tokens grouped by expert ([sum(map), h]) and
    the sorted source indices (row i of the output came from token ``sorted_indices[i]``)."""
    n, E = routing_map.shape
    rm = routing_map.bool().T.contiguous()            # [E, n]
    token_idx = torch.arange(n, device=tokens.device).unsqueeze(0).expand(E, -1)
    sorted_indices = token_idx.masked_select(rm)
    if num_out_tokens is not None and not drop_and_pad:
        sorted_indices = sorted_indices[:num_out_tokens]
    return tokens.index_select(0, sorted_indices), sorted_indices


def unpermute(permuted_tokens, sorted_indices, restore_shape, probs=None, routing_map=None):
    n, h = restore_shape
    if probs is not None:
        p = probs.T.contiguous().masked_select(routing_map.bool().T.contiguous())
        permuted_tokens = permuted_tokens * p.unsqueeze(-1).to(permuted_tokens.dtype)
    out = torch.zeros(n, h, dtype=permuted_tokens.dtype, device=permuted_tokens.device)
    out.index_add_(0, sorted_indices, permuted_tokens)
    return out


def sort_chunks_by_idxs(x, split_sizes, sorted_idxs):
    chunks = torch.split(x, split_sizes.tolist() if torch.is_tensor(split_sizes) else list(split_sizes), dim=0)
    return torch.cat([chunks[i] for i in (sorted_idxs.tolist() if torch.is_tensor(sorted_idxs) else sorted_idxs)], dim=0)


def group_limited_topk(scores, topk, num_tokens, num_experts, num_groups, group_topk):
    group_scores = scores.view(num_tokens, num_groups, -1).topk(max(1, topk // group_topk), dim=-1)[0].sum(dim=-1)
    group_idx = torch.topk(group_scores, k=group_topk, dim=-1, sorted=False)[1]
    group_mask = torch.zeros_like(group_scores).scatter_(1, group_idx, 1)
    score_mask = group_mask.unsqueeze(-1).expand(num_tokens, num_groups, num_experts // num_groups).reshape(num_tokens, -1)
    masked = scores.masked_fill(~score_mask.bool(), float("-inf"))
    return torch.topk(masked, k=topk, dim=-1)


def topk_softmax_with_capacity(logits, topk, capacity_factor=None, pad_to_capacity=False, drop_policy="probs",
                               use_pre_softmax=False, num_groups=None, group_topk=None, scaling_factor=None,
                               deterministic_mode=False, score_function="softmax", expert_bias=None):
    n, E = logits.shape

    def _topk(scores):
        if group_topk:
            return group_limited_topk(scores, topk, n, E, num_groups, group_topk)
        return torch.topk(scores, k=topk, dim=1)

    if score_function == "softmax":
        if use_pre_softmax:
            scores = torch.softmax(logits, dim=-1, dtype=torch.float32).type_as(logits)
            probs, top_idx = _topk(scores)
        else:
            sc, top_idx = _topk(logits)
            probs = torch.softmax(sc, dim=-1, dtype=torch.float32).type_as(logits)
    elif score_function == "sigmoid":
        scores = torch.sigmoid(logits.float()).type_as(logits)
        if expert_bias is not None:
            _, top_idx = _topk(scores + expert_bias)
            sc = torch.gather(scores, 1, top_idx)
        else:
            sc, top_idx = _topk(scores)
        probs = sc / (sc.sum(dim=-1, keepdim=True) + 1e-20) if topk > 1 else sc
    else:
        raise ValueError(f"invalid score_function {score_function}")
    if scaling_factor:
        probs = probs * scaling_factor
    topk_masked_gates = torch.zeros_like(logits).scatter(1, top_idx, probs)
    topk_map = torch.zeros_like(logits).int().scatter(1, top_idx, 1).bool()
    tokens_per_expert = topk_map.sum(dim=0)
    if capacity_factor is None:
        return topk_masked_gates, topk_map, tokens_per_expert
    cap = get_capacity(n * topk, E, capacity_factor)
    if drop_policy == "probs":
        _, cap_idx = torch.topk(topk_masked_gates, k=min(cap, n), dim=0, sorted=False)
        cap_mask = torch.zeros_like(logits).scatter(0, cap_idx, 1).bool()
    elif drop_policy == "position":
        _, cap_idx = torch.topk(topk_map.int(), k=min(cap, n), dim=0, sorted=False)
        cap_mask = torch.zeros_like(logits).scatter(0, cap_idx, 1).bool()
    else:
        raise ValueError(f"invalid drop_policy {drop_policy}")
    if pad_to_capacity:
        final_map, final_probs = cap_mask, topk_masked_gates * cap_mask
    else:
        final_map = topk_map & cap_mask
        final_probs = topk_masked_gates * final_map
    return final_probs, final_map, tokens_per_expert


def save_to_aux_losses_tracker(name, loss, layer_number, num_layers, reduce_group=None, avg_group=None):
    if layer_number is None:
        return
    tr = _MOE_LAYER_WISE_LOGGING_TRACKER
    if name not in tr:
        tr[name] = {"values": torch.zeros(num_layers, device=loss.device)}
    tr[name]["values"][layer_number - 1] += loss.detach()
    tr[name]["reduce_group"], tr[name]["avg_group"] = reduce_group, avg_group


def clear_aux_losses_tracker():
    for v in _MOE_LAYER_WISE_LOGGING_TRACKER.values():
        v["values"].zero_()


def get_moe_layer_wise_logging_tracker():
    return _MOE_LAYER_WISE_LOGGING_TRACKER


def reduce_aux_losses_tracker_across_ranks():
    """Layers live on different pipeline stages: sum the per-layer vectors over the PP group so that the logging rank
    sees every layer (reference ``moe_utils.reduce_aux_losses_tracker_across_ranks``), then average over the data-parallel
    replicas."""
    import torch.distributed as dist

    from ..parallel import state as ps

    if not (dist.is_available() and dist.is_initialized() and ps.model_parallel_is_initialized()):
        return
    for d in _MOE_LAYER_WISE_LOGGING_TRACKER.values():
        v = d["values"]
        for group, avg in ((ps.get_pipeline_model_parallel_group(), False), (ps.get_data_parallel_group(), True)):
            if group is not None and not isinstance(group, ps._SingleGroup):
                dist.all_reduce(v, group=group)
                if avg:
                    v.div_(dist.get_world_size(group))


def track_moe_metrics(loss_scale, iteration, writer, wandb_writer=None, total_loss_dict=None, per_layer_logging=False):
    reduce_aux_losses_tracker_across_ranks()
    for name, d in _MOE_LAYER_WISE_LOGGING_TRACKER.items():
        vals = d["values"].float() * loss_scale
        if total_loss_dict is not None:
            total_loss_dict[name] = total_loss_dict.get(name, 0.0) + float(vals.mean())
        if writer is not None:
            writer.add_scalar(name, float(vals.mean()), iteration)
            if per_layer_logging:
                for i, v in enumerate(vals.tolist()):
                    writer.add_scalar(f"moe/{name}_layer_{i}", v, iteration)
    clear_aux_losses_tracker()


def sinkhorn(cost, tol: float = 1e-4):
    from .router import sinkhorn as impl

    return impl(cost, tol)


def get_updated_expert_bias(tokens_per_expert, expert_bias, expert_bias_update_rate):
    """Auxiliary-loss-free load balancing (DeepSeek-V3): experts that received fewer tokens than the average get their routing
    bias raised by ``update_rate``, over-loaded ones lowered; token counts are summed over the whole data-parallel domain first."""
    import torch.distributed as dist

    from ..parallel import state as ps

    with torch.no_grad():
        counts = tokens_per_expert.float().clone()
        if ps.model_parallel_is_initialized() and dist.is_available() and dist.is_initialized():
            g = ps.get_tensor_and_data_parallel_group(with_context_parallel=True)
            if not type(g).__name__.startswith("_Single") and dist.get_world_size(g) > 1:
                dist.all_reduce(counts, group=g)
        offset = counts.mean(dim=-1, keepdim=True) - counts
        return expert_bias + torch.sign(offset) * expert_bias_update_rate


def maybe_move_tensor_to_cpu(tensor, as_numpy: bool = False, record_stream: bool = False):
    """Asynchronous device -> pinned host copy of a (small) routing tensor; the caller synchronises before reading it."""
    if torch.is_tensor(tensor) and tensor.is_cuda:
        out = torch.empty(tensor.size(), dtype=tensor.dtype, device="cpu", pin_memory=True)
        out.copy_(tensor, non_blocking=True)
        if record_stream:
            tensor.record_stream(torch.cuda.current_stream())
        tensor = out
    return tensor.numpy() if (as_numpy and torch.is_tensor(tensor)) else tensor
