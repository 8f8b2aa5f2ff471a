"""Sharded ("distributed") checkpoints with resharding on load.

Capability parity with reference ``core/dist_checkpointing`` (``ShardedTensor`` metadata
``mapping.py:79-89``: key, local/global shape, global offset, axis fragmentation, replica id,
``allow_shape_mismatch``; ``save`` / ``load`` in ``serialization.py``; ``common.pt`` for
non-tensor state ``strategies/common.py``; TP/PP/DP resharding on load as a *format* feature;
fully-parallel save = every rank writes only the shards it owns with ``replica_id == 0``).

On-disk layout of one checkpoint directory::

    common.pt                      non-tensor state (args, iteration, rng, scheduler ...)
    metadata.json                  {key: {global_shape, dtype, shards: [{file, offset, shape}]}}
    shards_{rank:05d}.pt           {"<key>@<offset tuple>": tensor}  (one file per writing rank)

Loading intersects each requested local shard (its global offset/shape under the *new*
parallel layout) with the stored shards, so a checkpoint written at TP4 loads at TP2/TP8, PP
changes only rename layer prefixes (handled by the caller's ``sharded_state_dict``).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist

from ..parallel import state as ps


@dataclass
class ShardedTensor:
    key: str
    data: Optional[torch.Tensor]
    dtype: torch.dtype
    local_shape: Tuple[int, ...]
    global_shape: Tuple[int, ...]
    global_offset: Tuple[int, ...]
    axis_fragmentations: Optional[Tuple[int, ...]] = None
    replica_id: Tuple[int, ...] = (0,)
    prepend_axis_num: int = 0
    allow_shape_mismatch: bool = False
    flattened_range: Optional[slice] = None
    replica_count: Optional[int] = None   # size of the group the last replica_id component indexes (fully-parallel save)

    @classmethod
    def from_rank_offsets(cls, key, data, *rank_offsets, replica_id=(0,), prepend_axis_num=0,
                          allow_shape_mismatch=False):
        """``rank_offsets``: tuples ``(axis, rank_along_axis, fragments_along_axis)``."""
        gshape = list(data.shape)
        goff = [0] * data.dim()
        frag = [1] * data.dim()
        for axis, r, n in rank_offsets:
            gshape[axis] = data.shape[axis] * n
            goff[axis] = data.shape[axis] * r
            frag[axis] = n
        return cls(key, data, data.dtype, tuple(data.shape), tuple(gshape), tuple(goff), tuple(frag),
                   replica_id, prepend_axis_num, allow_shape_mismatch)

    def is_main_replica(self) -> bool:
        r = self.replica_id
        return r == 0 or (isinstance(r, (tuple, list)) and all(x == 0 for x in r))


@dataclass
class ShardedObject:
    key: str
    data: object
    replica_id: Tuple[int, ...] = (0,)
    global_shape: Tuple[int, ...] = (1,)
    global_offset: Tuple[int, ...] = (0,)


def _dp_world() -> int:
    try:
        return ps.get_data_parallel_world_size(with_context_parallel=True) if ps.model_parallel_is_initialized() else 1
    except Exception:  # noqa: BLE001
        return 1


def _from_offsets_with_prepend(key, data, prepend_offsets, data_offsets, replica_id, allow_shape_mismatch=False,
                               prepend_axis_num=None, **_kw):
    """``prepend_offsets`` (axis, rank, fragments) describe axes that exist only in the GLOBAL tensor (stacked layers /
    experts, reference ``ShardedTensor.from_rank_offsets(..., prepend_axis_num=)``); ``data_offsets`` index axes of ``data``."""
    prepend_offsets = tuple(prepend_offsets)
    pre = len(prepend_offsets)
    gshape = [1] * pre + list(data.shape)
    goff = [0] * (pre + data.dim())
    frag = [1] * (pre + data.dim())
    for i, (axis, r, n) in enumerate(prepend_offsets):
        gshape[i], goff[i], frag[i] = n, r, n
    for axis, r, n in data_offsets:
        a = axis + pre
        gshape[a] = data.shape[axis] * n
        goff[a] = data.shape[axis] * r
        frag[a] = n
    return ShardedTensor(key, data, data.dtype, tuple(data.shape), tuple(gshape), tuple(goff), tuple(frag), replica_id,
                         pre, allow_shape_mismatch)


def make_tp_sharded_tensor_for_checkpoint(tensor, key, tp_axis=0, replica_id=None, prepend_offsets=(), **kw):
    tp_rank, tp_size = ps.get_tensor_model_parallel_rank(), ps.get_tensor_model_parallel_world_size()
    dp_rank = ps.get_data_parallel_rank(with_context_parallel=True) if ps.model_parallel_is_initialized() else 0
    if replica_id is None:
        replica_id = (0, 0, dp_rank)
    st = _from_offsets_with_prepend(key, tensor, prepend_offsets, [(tp_axis, tp_rank, tp_size)], replica_id, **kw)
    st.replica_count = _dp_world()
    return st


def make_sharded_tensor_for_checkpoint(tensor, key, prepend_offsets=(), replica_id=None, **kw):
    tp_rank = ps.get_tensor_model_parallel_rank()
    dp_rank = ps.get_data_parallel_rank(with_context_parallel=True) if ps.model_parallel_is_initialized() else 0
    if replica_id is None:
        replica_id = (0, tp_rank, dp_rank)
    st = _from_offsets_with_prepend(key, tensor, prepend_offsets, [], replica_id, **kw)
    st.replica_count = _dp_world()
    return st


def make_sharded_tensors_for_checkpoint(state_dict: Dict[str, torch.Tensor], prefix: str,
                                        tensor_parallel_layers_axis_map: Optional[Dict[str, int]] = None,
                                        sharded_offsets: Iterable[Tuple[int, int, int]] = (),
                                        allow_shape_mismatch: Iterable[str] = ()):
    amap = tensor_parallel_layers_axis_map or {}
    out = {}
    for name, t in state_dict.items():
        if t is None or name.endswith("_extra_state"):
            continue
        key = f"{prefix}{name}"
        if name in amap and getattr(t, "tensor_model_parallel", True):
            out[key] = make_tp_sharded_tensor_for_checkpoint(t, key, amap[name], prepend_offsets=tuple(sharded_offsets),
                                                             allow_shape_mismatch=name in allow_shape_mismatch)
        else:
            out[key] = make_sharded_tensor_for_checkpoint(t, key, prepend_offsets=tuple(sharded_offsets))
    return out


# ---------------------------------------------------------------------------
# save / load
# ---------------------------------------------------------------------------

def _rank():
    return dist.get_rank() if dist.is_initialized() else 0


def _world():
    return dist.get_world_size() if dist.is_initialized() else 1


def _dtype_str(dt) -> str:
    return str(dt).replace("torch.", "")


def save(sharded_state_dict: Dict[str, object], checkpoint_dir: str, common_state: Optional[dict] = None) -> None:
    os.makedirs(checkpoint_dir, exist_ok=True)
    rank = _rank()
    mine, meta = {}, {}
    for k, v in sharded_state_dict.items():
        if isinstance(v, ShardedTensor):
            data = v.data
            if hasattr(data, "full_tensor"):          # torch FSDP2 DTensor: assembling it is a collective -> every rank
                data = data.full_tensor()
            if not v.is_main_replica():
                continue
            sk = f"{v.key}@{','.join(map(str, v.global_offset))}"
            mine[sk] = data.detach().to("cpu", copy=True)
            meta.setdefault(v.key, {"global_shape": list(v.global_shape), "dtype": _dtype_str(v.dtype), "shards": []})
            meta[v.key]["shards"].append({"file": f"shards_{rank:05d}.pt", "offset": list(v.global_offset),
                                          "shape": list(v.local_shape), "name": sk})
    torch.save(mine, os.path.join(checkpoint_dir, f"shards_{rank:05d}.pt"))
    if dist.is_initialized() and _world() > 1:
        metas = [None] * _world()
        dist.all_gather_object(metas, meta)
    else:
        metas = [meta]
    if rank == 0:
        merged: Dict[str, dict] = {}
        for m in metas:
            for key, info in m.items():
                if key not in merged:
                    merged[key] = {"global_shape": info["global_shape"], "dtype": info["dtype"], "shards": []}
                merged[key]["shards"].extend(info["shards"])
        with open(os.path.join(checkpoint_dir, "metadata.json"), "w") as f:
            json.dump(merged, f)
        if common_state is not None:
            torch.save(common_state, os.path.join(checkpoint_dir, "common.pt"))
    if dist.is_initialized() and _world() > 1:
        dist.barrier()


def load_common_state_dict(checkpoint_dir: str) -> dict:
    p = os.path.join(checkpoint_dir, "common.pt")
    return torch.load(p, map_location="cpu", weights_only=False) if os.path.exists(p) else {}


def load(sharded_state_dict: Dict[str, object], checkpoint_dir: str, strict: bool = True) -> Dict[str, torch.Tensor]:
    """Fill every ShardedTensor's ``data`` in place from the stored shards; returns {key: tensor}."""
    with open(os.path.join(checkpoint_dir, "metadata.json")) as f:
        meta = json.load(f)
    cache: Dict[str, dict] = {}

    def shard_file(name):
        if name not in cache:
            cache[name] = torch.load(os.path.join(checkpoint_dir, name), map_location="cpu", weights_only=False)
        return cache[name]

    out, missing = {}, []
    for k, v in sharded_state_dict.items():
        if not isinstance(v, ShardedTensor):
            continue
        info = meta.get(v.key)
        if info is None:
            missing.append(v.key)
            continue
        gshape = tuple(info["global_shape"])
        if gshape != tuple(v.global_shape) and not v.allow_shape_mismatch:
            raise ValueError(f"global shape mismatch for {v.key}: checkpoint {gshape} vs model {v.global_shape}")
        dst = torch.zeros(v.local_shape, dtype=v.dtype)
        lo = v.global_offset
        hi = tuple(o + s for o, s in zip(lo, v.local_shape))
        covered = 0
        for sh in info["shards"]:
            so, ss = sh["offset"], sh["shape"]
            se = [o + s for o, s in zip(so, ss)]
            inter_lo = [max(a, b) for a, b in zip(lo, so)]
            inter_hi = [min(a, b) for a, b in zip(hi, se)]
            if any(a >= b for a, b in zip(inter_lo, inter_hi)):
                continue
            src = shard_file(sh["file"])[sh["name"]]
            src_sl = tuple(slice(a - o, b - o) for a, b, o in zip(inter_lo, inter_hi, so))
            dst_sl = tuple(slice(a - o, b - o) for a, b, o in zip(inter_lo, inter_hi, lo))
            dst[dst_sl] = src[src_sl].to(v.dtype)
            n = 1
            for a, b in zip(inter_lo, inter_hi):
                n *= b - a
            covered += n
        if covered < dst.numel() and not v.allow_shape_mismatch:
            raise ValueError(f"checkpoint shards do not cover {v.key} (covered {covered} of {dst.numel()})")
        if v.data is not None and hasattr(v.data, "full_tensor"):
            from torch.distributed.tensor import distribute_tensor

            local = distribute_tensor(dst.to(v.data.device), v.data.device_mesh, v.data.placements)
            with torch.no_grad():
                v.data.detach().to_local().copy_(local.to_local())
        elif v.data is not None:
            v.data.data.copy_(dst.to(v.data.device))
        out[k] = dst
    if missing and strict:
        raise KeyError(f"keys missing from the checkpoint: {missing[:8]}{'...' if len(missing) > 8 else ''}")
    return out
