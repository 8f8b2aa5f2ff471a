"""``megatron.core.tensor_parallel.utils``."""
from typing import List, Sequence

import torch
import torch.distributed as dist

from ..parallel import state as ps
from ..parallel.tp_layers import VocabUtility  # noqa: F401


def split_tensor_along_last_dim(tensor: torch.Tensor, num_partitions: int, contiguous_split_chunks: bool = False) -> List[torch.Tensor]:
    if tensor.size(-1) % num_partitions != 0:
        raise AssertionError(f"{tensor.size(-1)} is not divisible by {num_partitions}")
    parts = torch.split(tensor, tensor.size(-1) // num_partitions, dim=-1)
    return [p.contiguous() for p in parts] if contiguous_split_chunks else list(parts)


def split_tensor_into_1d_equal_chunks(tensor: torch.Tensor, new_buffer: bool = False) -> torch.Tensor:
    """This tensor-parallel rank's equal slice of the flattened tensor."""
    world, rank = ps.get_tensor_model_parallel_world_size(), ps.get_tensor_model_parallel_rank()
    n = tensor.numel() // world
    flat = tensor.reshape(-1)[rank * n:(rank + 1) * n]
    return flat.clone() if new_buffer else flat


def gather_split_1d_tensor(tensor: torch.Tensor) -> torch.Tensor:
    """Inverse of ``split_tensor_into_1d_equal_chunks``."""
    world = ps.get_tensor_model_parallel_world_size()
    if world == 1:
        return tensor.reshape(-1)
    out = torch.empty(world * tensor.numel(), dtype=tensor.dtype, device=tensor.device)
    dist.all_gather_into_tensor(out, tensor.reshape(-1).contiguous(), group=ps.get_tensor_model_parallel_group())
    return out
