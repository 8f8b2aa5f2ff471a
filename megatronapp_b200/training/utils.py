"""Harness helpers (reference ``training/utils.py``): rank-aware printing, batch broadcast over
the TP group, parameter norms, memory report."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ..parallel import state as ps


def unwrap_model(model, module_instances=None):
    is_list = isinstance(model, list)
    ms = model if is_list else [model]
    out = []
    for m in ms:
        while hasattr(m, "module"):
            m = m.module
        out.append(m)
    return out if is_list else out[0]


def is_last_rank():
    return (not dist.is_initialized()) or dist.get_rank() == dist.get_world_size() - 1


def print_rank_0(msg):
    if not dist.is_initialized() or dist.get_rank() == 0:
        print(msg, flush=True)


def print_rank_last(msg):
    if is_last_rank():
        print(msg, flush=True)


def report_memory(name):
    if not torch.cuda.is_available():
        return
    mb = 1024.0 * 1024.0
    s = f"memory (MB) | allocated: {torch.cuda.memory_allocated() / mb:.1f} | max allocated: " \
        f"{torch.cuda.max_memory_allocated() / mb:.1f} | reserved: {torch.cuda.memory_reserved() / mb:.1f} | " \
        f"max reserved: {torch.cuda.max_memory_reserved() / mb:.1f}"
    if ps.get_data_parallel_rank() == 0:
        print(f"[Rank {dist.get_rank() if dist.is_initialized() else 0}] {name} {s}", flush=True)


def _real(g):
    return g is not None and not isinstance(g, ps._SingleGroup) and dist.is_initialized() and dist.get_world_size(g) > 1


def _dev():
    return "cuda" if (torch.cuda.is_available() and dist.is_initialized() and dist.get_backend() == "nccl") \
        else ("cuda" if (torch.cuda.is_available() and not dist.is_initialized()) else "cpu")


def logical_and_across_model_parallel_group(flag: bool) -> bool:
    g = ps.get_model_parallel_group(check_initialized=False) if ps.model_parallel_is_initialized() else None
    if not _real(g):
        return bool(flag)
    t = torch.tensor([int(bool(flag))], dtype=torch.int, device=_dev())
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=g)
    return bool(t.item())


def reduce_max_stat_across_model_parallel_group(stat):
    g = ps.get_model_parallel_group(check_initialized=False) if ps.model_parallel_is_initialized() else None
    if not _real(g):
        return stat
    t = torch.tensor([-1.0 if stat is None else float(stat)], dtype=torch.float32, device=_dev())
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=g)
    return None if t.item() == -1.0 else t.item()


def calc_params_l2_norm(model):
    from .. import ops

    ms = model if isinstance(model, list) else [model]
    ts = []
    for m in ms:
        for p in m.parameters():
            dup = (not getattr(p, "tensor_model_parallel", False)) and ps.get_tensor_model_parallel_rank() != 0
            if not dup and not getattr(p, "shared", False):
                ts.append(p.data.reshape(-1))
    sq = ops.l2norm_sq(ts)
    g = ps.get_model_parallel_group(check_initialized=False) if ps.model_parallel_is_initialized() else None
    if _real(g):
        dist.all_reduce(sq, group=g)
    return float(sq.sqrt().item())


def average_losses_across_data_parallel_group(losses):
    t = torch.cat([l.clone().detach().view(1) for l in losses])
    g = ps.get_data_parallel_group()
    if _real(g):
        dist.all_reduce(t, group=g)
        t = t / dist.get_world_size(g)
    return t


def get_batch_on_this_tp_rank(data_iterator, args, pinned_stage=None):
    """TP-rank 0 loads the microbatch and broadcasts it to its TP group (reference
    ``training/utils.py get_batch_on_this_tp_rank``); first stage needs tokens/position ids, last
    stage labels/loss mask."""
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    tp_group = ps.get_tensor_model_parallel_group()
    multi = _real(tp_group)
    src = ps.get_tensor_model_parallel_src_rank() if multi else 0
    keys = ["tokens", "labels", "loss_mask", "position_ids"]
    if args.create_attention_mask_in_dataloader:
        keys.append("attention_mask")
    if ps.get_tensor_model_parallel_rank() == 0:
        data = next(data_iterator) if data_iterator is not None else None
        batch = {k: (data[k].to(dev, non_blocking=True) if (data is not None and k in data) else None) for k in keys}
        if "attention_mask" not in batch:
            batch["attention_mask"] = None
        if multi:
            for k in keys:
                if batch.get(k) is not None:
                    dist.broadcast(batch[k], src, group=tp_group)
    else:
        b, s = args.micro_batch_size, args.seq_length
        batch = {"tokens": torch.empty((b, s), dtype=torch.int64, device=dev),
                 "labels": torch.empty((b, s), dtype=torch.int64, device=dev),
                 "loss_mask": torch.empty((b, s), dtype=torch.float32, device=dev),
                 "position_ids": torch.empty((b, s), dtype=torch.int64, device=dev),
                 "attention_mask": (torch.empty((b, 1, s, s), dtype=torch.bool, device=dev)
                                    if args.create_attention_mask_in_dataloader else None)}
        for k in keys:
            if batch.get(k) is not None:
                dist.broadcast(batch[k], src, group=tp_group)
    return batch


def broadcast_batch(data_iterator, int_keys, float_keys=()):
    """Fetch the next micro-batch on TP rank 0 and broadcast it over the TP group (the other TP ranks hold no data
    iterator) -- the ``tensor_parallel.broadcast_data`` idiom of the reference's ``pretrain_bert / pretrain_t5``."""
    from ..parallel import state as ps
    from ..parallel.data import broadcast_data

    data = None
    if data_iterator is not None and ps.get_tensor_model_parallel_rank() == 0:
        raw = next(data_iterator)
        data = {k: torch.as_tensor(raw[k]) for k in list(int_keys) + list(float_keys)}
        data = {k: (v.long() if k in int_keys else v.float()) for k, v in data.items()}
    out = broadcast_data(list(int_keys), data, torch.int64)
    if float_keys:
        out.update(broadcast_data(list(float_keys), data, torch.float32))
    return out


# -- further helpers of the reference's ``megatron/training/utils.py`` ------------------------------------------------------
def is_rank0() -> bool:
    return (not dist.is_initialized()) or dist.get_rank() == 0


def get_ltor_masks_and_position_ids(data, eod_token, reset_position_ids=False, reset_attention_mask=False, eod_mask_loss=False, **kw):
    """Left-to-right attention mask, loss mask and position ids of a ``[b, s]`` token batch (document boundaries at ``eod_token``)."""
    from ..data.mock import get_ltor_masks_and_position_ids as impl

    return impl(data, eod_token, reset_position_ids, reset_attention_mask, eod_mask_loss, **kw)


def append_to_progress_log(string: str, barrier: bool = True):
    from .training import append_to_progress_log as impl

    return impl(string, barrier)


def get_blend_and_blend_per_split(args):
    """``(blend, blend_per_split)`` in the reference's shape: ``(prefixes, weights)`` tuples or ``None``."""
    from ..data.gpt_dataset import blend_and_blend_per_split, parse_blend

    blend, per_split = blend_and_blend_per_split(args)
    as_tuple = lambda b: None if not b else parse_blend(b)   # noqa: E731
    return as_tuple(blend), (None if per_split is None else [as_tuple(b) for b in per_split])


def print_params_min_max_norm(optimizer, iteration):
    """One line per parameter: iteration, rank, index, model-parallel flag, min, max, L2 norm."""
    lines, index = "iteration, rank, index, tensor-model-parallel, min, max, norm\n", 0
    rank = dist.get_rank() if dist.is_initialized() else 0
    for group in optimizer.param_groups:
        for p in group["params"]:
            index += 1
            lines += (f"{iteration:7d}, {rank:4d}, {index:4d}, {int(getattr(p, 'tensor_model_parallel', False)):2d}, "
                      f"{float(p.data.min()):.6E}, {float(p.data.max()):.6E}, {float(p.data.float().norm()):.6E}\n")
    print(lines, flush=True)


def check_adlr_autoresume_termination(iteration, model, optimizer, opt_param_scheduler):
    """Checkpoint and exit when the cluster's auto-resume hook asks for termination."""
    from .autoresume import check_adlr_autoresume_termination as impl

    return impl(iteration, model, optimizer, opt_param_scheduler)


def get_device_arch_version() -> int:
    """Major compute capability of the current GPU (10 on B200)."""
    return torch.cuda.get_device_properties(torch.cuda.current_device()).major if torch.cuda.is_available() else 0


def update_use_dist_ckpt(args):
    args.use_dist_ckpt = args.ckpt_format != "torch"
