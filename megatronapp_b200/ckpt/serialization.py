"""``dist_checkpointing`` entry points of the reference (``core/dist_checkpointing/{__init__,serialization,validation,core}.py``)
over this framework's ``torch_dist`` reader / writer: ``save`` / ``load`` with strictness modes, whole-tensor and metadata-only
loads, removal of a key family (e.g. the optimizer state) from an existing checkpoint, and validation that the shards of a
sharded state dict tile every global tensor exactly once."""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass
from enum import Enum
from typing import Dict, Iterable, List, Optional, Set, Tuple

import torch
import torch.distributed as dist

from . import torch_dist
from .sharded import ShardedObject, ShardedTensor


class CheckpointingException(Exception):
    pass


class StrictHandling(Enum):
    """What ``load`` does about keys that are requested but absent ("missing") or present but not requested ("unexpected")."""
    ASSUME_OK_UNEXPECTED = "assume_ok_unexpected"     # no checks at all (a missing key surfaces as a KeyError while loading)
    LOG_UNEXPECTED = "log_unexpected"                 # unexpected keys are printed
    LOG_ALL = "log_all"                               # missing and unexpected keys are printed, loading continues
    RAISE_UNEXPECTED = "raise_unexpected"
    RAISE_ALL = "raise_all"
    RETURN_UNEXPECTED = "return_unexpected"           # load returns (state_dict, missing (always empty), unexpected)
    RETURN_ALL = "return_all"                         # load returns (state_dict, missing, unexpected)
    IGNORE_ALL = "ignore_all"


@dataclass
class LocalNonpersistentObject:
    """An object that takes part in the state dict but is never written: ``load`` hands the local value back untouched
    (reference ``mapping.LocalNonpersistentObject``: e.g. RNG states kept per rank)."""
    obj: object

    def unwrap(self):
        return self.obj


def check_is_distributed_checkpoint(checkpoint_dir: str) -> bool:
    return os.path.isfile(os.path.join(checkpoint_dir, "metadata.json"))


def save(sharded_state_dict: Dict[str, object], checkpoint_dir: str, common_state: Optional[dict] = None, async_save: bool = False,
         validate_access_integrity: bool = True):
    persistent = {k: v for k, v in sharded_state_dict.items() if not isinstance(v, LocalNonpersistentObject)}
    if validate_access_integrity:
        validate_sharding_integrity(persistent)
    return torch_dist.save(persistent, checkpoint_dir, common_state, async_save=async_save)


def _requested_keys(ssd: Dict[str, object]) -> Set[str]:
    out = set()
    for v in ssd.values():
        if isinstance(v, ShardedTensor):
            out.add(v.key)
        elif isinstance(v, ShardedObject):
            out.add(torch_dist.sharded_object_key(v))
    return out


def load(sharded_state_dict: Dict[str, object], checkpoint_dir: str, strict=StrictHandling.ASSUME_OK_UNEXPECTED):
    strict = StrictHandling(strict) if not isinstance(strict, StrictHandling) else strict
    local = {k: v.unwrap() for k, v in sharded_state_dict.items() if isinstance(v, LocalNonpersistentObject)}
    ssd = {k: v for k, v in sharded_state_dict.items() if not isinstance(v, LocalNonpersistentObject)}
    md = torch_dist.read_metadata(checkpoint_dir)
    have = set(md.state_dict_metadata)
    want = _requested_keys(ssd)
    missing, unexpected = sorted(want - have), sorted(have - want)
    if strict in (StrictHandling.RAISE_ALL,) and missing:
        raise CheckpointingException(f"missing keys: {missing[:8]}")
    if strict in (StrictHandling.RAISE_ALL, StrictHandling.RAISE_UNEXPECTED) and unexpected:
        raise CheckpointingException(f"unexpected keys in the checkpoint: {unexpected[:8]}")
    if strict in (StrictHandling.LOG_ALL, StrictHandling.LOG_UNEXPECTED) and unexpected and (not dist.is_initialized() or dist.get_rank() == 0):
        print(f"WARNING: {len(unexpected)} checkpoint keys are not requested, e.g. {unexpected[:4]}", flush=True)
    if strict is StrictHandling.LOG_ALL and missing and (not dist.is_initialized() or dist.get_rank() == 0):
        print(f"WARNING: {len(missing)} requested keys are not in the checkpoint, e.g. {missing[:4]}", flush=True)
    tolerate_missing = strict in (StrictHandling.LOG_ALL, StrictHandling.RETURN_ALL, StrictHandling.IGNORE_ALL)
    if tolerate_missing and missing:
        gone = set(missing)
        ssd = {k: v for k, v in ssd.items() if not ((isinstance(v, ShardedTensor) and v.key in gone)
                                                   or (isinstance(v, ShardedObject) and torch_dist.sharded_object_key(v) in gone))}
    out = torch_dist.load(ssd, checkpoint_dir, strict=not tolerate_missing)
    out.update(local)
    if strict is StrictHandling.RETURN_ALL:
        return out, missing, unexpected
    if strict is StrictHandling.RETURN_UNEXPECTED:
        return out, [], unexpected
    return out


def load_common_state_dict(checkpoint_dir: str) -> dict:
    return torch_dist.load_common_state_dict(checkpoint_dir)


def load_tensors_metadata(checkpoint_dir: str) -> Dict[str, ShardedTensor]:
    """Every tensor of the checkpoint as a data-less ShardedTensor covering the whole global shape."""
    from torch.distributed.checkpoint.metadata import TensorStorageMetadata

    out = {}
    for k, m in torch_dist.read_metadata(checkpoint_dir).state_dict_metadata.items():
        if isinstance(m, TensorStorageMetadata):
            shape = tuple(int(x) for x in m.size)
            out[k] = ShardedTensor(k, None, m.properties.dtype, shape, shape, (0,) * len(shape), (1,) * len(shape))
    return out


def load_sharded_metadata(checkpoint_dir: str) -> Dict[str, object]:
    """Tensors as above plus the sharded objects' keys (reference ``load_sharded_metadata``)."""
    from torch.distributed.checkpoint.metadata import BytesStorageMetadata

    out: Dict[str, object] = dict(load_tensors_metadata(checkpoint_dir))
    for k, m in torch_dist.read_metadata(checkpoint_dir).state_dict_metadata.items():
        if isinstance(m, BytesStorageMetadata):
            out[k] = ShardedObject(k, None)
    return out


def load_plain_tensors(checkpoint_dir: str) -> Dict[str, torch.Tensor]:
    """Whole (unsharded) tensors by key: converters and inspection tools."""
    ssd = load_tensors_metadata(checkpoint_dir)
    return {k: v for k, v in torch_dist.load(ssd, checkpoint_dir, strict=True).items() if torch.is_tensor(v)}


def remove_sharded_tensors(checkpoint_dir: str, key_prefix: str) -> List[str]:
    """Drop every key starting with ``key_prefix`` from the checkpoint's metadata (e.g. ``optimizer.`` to turn a training
    checkpoint into a release checkpoint).  The data files keep the bytes; they are simply no longer referenced."""
    path = os.path.join(checkpoint_dir, ".metadata")
    md = torch_dist.read_metadata(checkpoint_dir)
    gone = [k for k in md.state_dict_metadata if k.startswith(key_prefix)]
    for k in gone:
        del md.state_dict_metadata[k]
    if md.storage_data:
        for idx in [i for i in md.storage_data if i.fqn.startswith(key_prefix)]:
            del md.storage_data[idx]
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        pickle.dump(md, f)
    os.replace(tmp, path)
    return gone


# ---- validation --------------------------------------------------------------------------------------------------------

def _describe(ssd: Dict[str, object]) -> List[Tuple]:
    rows = []
    for v in ssd.values():
        if isinstance(v, ShardedTensor):
            rows.append((v.key, tuple(int(x) for x in v.global_shape), tuple(int(x) for x in v.global_offset),
                         (1,) * v.prepend_axis_num + tuple(int(x) for x in v.local_shape), bool(v.is_main_replica()),
                         None if v.flattened_range is None else (int(v.flattened_range.start), int(v.flattened_range.stop))))
    return rows


def validate_sharding_integrity(sharded_state_dict: Dict[str, object], rows: Optional[Iterable[Tuple]] = None) -> None:
    """The main replicas of every key must cover its global shape exactly once: no gap, no overlap, no out-of-bounds shard
    (reference ``validation.validate_sharding_integrity``).  Collective when torch.distributed is initialised (every rank
    contributes the description of its shards); ``rows`` lets a test supply the descriptions of all ranks directly."""
    if rows is None:
        mine = _describe(sharded_state_dict)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            gathered: List[Optional[List[Tuple]]] = [None] * dist.get_world_size()
            dist.all_gather_object(gathered, mine)
            rows = [r for g in gathered for r in g]
        else:
            rows = mine
    by_key: Dict[str, List[Tuple]] = {}
    for r in rows:
        by_key.setdefault(r[0], []).append(r)
    for key, shards in by_key.items():
        gshape = shards[0][1]
        total = 1
        for x in gshape:
            total *= x
        covered = 0
        boxes = []
        for _, gs, off, shape, main, flat in shards:
            if gs != gshape:
                raise CheckpointingException(f"{key}: ranks disagree on the global shape ({gs} vs {gshape})")
            if any(o < 0 or o + s > g for o, s, g in zip(off, shape, gshape)):
                raise CheckpointingException(f"{key}: shard at {off} with shape {shape} leaves the global shape {gshape}")
            if not main:
                continue
            n = 1
            for x in shape:
                n *= x
            if flat is not None:
                if not (0 <= flat[0] <= flat[1] <= n):
                    raise CheckpointingException(f"{key}: flattened range {flat} outside the local shard of {n} elements")
                covered += flat[1] - flat[0]
                boxes.append((off, shape, flat))
            else:
                covered += n
                boxes.append((off, shape, None))
        for i in range(len(boxes)):
            for j in range(i + 1, len(boxes)):
                (o1, s1, f1), (o2, s2, f2) = boxes[i], boxes[j]
                if all(a < b + t and b < a + s for a, s, b, t in zip(o1, s1, o2, s2)):
                    if f1 is not None and f2 is not None and o1 == o2 and s1 == s2 and (f1[1] <= f2[0] or f2[1] <= f1[0]):
                        continue                                  # disjoint runs of the same flattened shard
                    raise CheckpointingException(f"{key}: main-replica shards overlap ({o1}, {s1}) and ({o2}, {s2})")
        if covered != total:
            raise CheckpointingException(f"{key}: main replicas cover {covered} of {total} elements")
