"""``megatron.core.transformer.utils``."""
from __future__ import annotations

from functools import lru_cache
from typing import Optional

import torch

from ..ckpt.sharded import ShardedObject, make_sharded_tensors_for_checkpoint  # noqa: F401
from ..transformer.mlp import erf_gelu  # noqa: F401


def get_linear_layer(rows: int, columns: int, init_method, perform_initialization: bool = True) -> torch.nn.Linear:
    layer = torch.nn.Linear(rows, columns)
    if perform_initialization:
        init_method(layer.weight)
    with torch.no_grad():
        layer.bias.zero_()
    return layer


@lru_cache(maxsize=32)
def get_default_causal_mask(sq: int) -> torch.Tensor:
    """Upper-triangular boolean mask (True = masked out)."""
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    return torch.triu(torch.ones(sq, sq, device=dev), diagonal=1).bool()


def attention_mask_func(attention_scores: torch.Tensor, attention_mask: Optional[torch.Tensor]) -> torch.Tensor:
    return attention_scores if attention_mask is None else attention_scores.masked_fill(attention_mask, -10000.0)


def gelu_impl(x):
    """tanh form of GELU (what the fused bias-GELU epilogue computes)."""
    return 0.5 * x * (1.0 + torch.tanh(0.7978845608028654 * x * (1.0 + 0.044715 * x * x)))


def openai_gelu(x):
    return gelu_impl(x)


def make_sharded_object_for_checkpoint(obj, key: str, sharded_offsets=(), replica_id=None, **kwargs) -> ShardedObject:
    from ..parallel import state as ps

    if replica_id is None:
        replica_id = (0, ps.get_tensor_model_parallel_rank(), ps.get_data_parallel_rank(with_context_parallel=True))
    offs = tuple(o[1] for o in sharded_offsets) if sharded_offsets else ()
    shape = tuple(o[2] for o in sharded_offsets) if sharded_offsets else ()
    return ShardedObject(key, obj, replica_id, global_shape=shape or (1,), global_offset=offs or (0,))


def sharded_state_dict_default(module: torch.nn.Module, prefix: str = "", sharded_offsets=(), metadata: Optional[dict] = None) -> dict:
    """``module.sharded_state_dict`` when it has one, else its plain tensors as replicated ShardedTensors."""
    if hasattr(module, "sharded_state_dict"):
        return module.sharded_state_dict(prefix=prefix)
    sd = {}
    module._save_to_state_dict(sd, "", keep_vars=True)
    return make_sharded_tensors_for_checkpoint(sd, prefix, {}, sharded_offsets)
