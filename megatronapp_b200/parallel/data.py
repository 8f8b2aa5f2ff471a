"""Broadcast a dict of tensors from TP-rank 0 to its tensor-parallel group with one size header and
one flat payload (reference ``tensor_parallel/data.py broadcast_data`` :106)."""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.distributed as dist

from . import state as ps

_MAX_DATA_DIM = 5


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if (torch.cuda.is_available() and dist.is_initialized()
                                                                 and dist.get_backend() == "nccl") else torch.device("cpu")


def _check_data_types(keys, data, target_dtype):
    for k in keys:
        if data[k].dtype != target_dtype:
            raise TypeError(f"{k} has data type {data[k].dtype} which is different than {target_dtype}")


def _build_key_size_numel_dictionaries(keys: List[str], data):
    sizes = [0] * (_MAX_DATA_DIM * len(keys))
    if ps.get_tensor_model_parallel_rank() == 0:
        off = 0
        for k in keys:
            assert data[k].dim() < _MAX_DATA_DIM, "you should increase MAX_DATA_DIM"
            for i, s in enumerate(data[k].size()):
                sizes[off + i] = s
            off += _MAX_DATA_DIM
    t = torch.tensor(sizes, dtype=torch.long, device=_dev())
    g = ps.get_tensor_model_parallel_group()
    if ps.get_tensor_model_parallel_world_size() > 1:
        dist.broadcast(t, ps.get_tensor_model_parallel_src_rank(), group=g)
    t = t.cpu().tolist()
    key_size, key_numel, total = {}, {}, 0
    off = 0
    for k in keys:
        shape, n = [], 1
        i = 0
        while i < _MAX_DATA_DIM and t[off + i] > 0:
            shape.append(t[off + i])
            n *= t[off + i]
            i += 1
        key_size[k], key_numel[k] = shape, n
        total += n
        off += _MAX_DATA_DIM
    return key_size, key_numel, total


def broadcast_data(keys: List[str], data: Dict[str, torch.Tensor], datatype: torch.dtype) -> Dict[str, torch.Tensor]:
    key_size, key_numel, total = _build_key_size_numel_dictionaries(keys, data)
    if ps.get_tensor_model_parallel_rank() == 0:
        _check_data_types(keys, data, datatype)
        flat = torch.cat([data[k].contiguous().view(-1) for k in keys], dim=0).to(_dev())
    else:
        flat = torch.empty(total, device=_dev(), dtype=datatype)
    if ps.get_tensor_model_parallel_world_size() > 1:
        dist.broadcast(flat, ps.get_tensor_model_parallel_src_rank(), group=ps.get_tensor_model_parallel_group())
    out, off = {}, 0
    for k in keys:
        out[k] = flat.narrow(0, off, key_numel[k]).view(key_size[k])
        off += key_numel[k]
    return out
