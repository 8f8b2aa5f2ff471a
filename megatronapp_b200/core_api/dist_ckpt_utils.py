"""Nested state-dict helpers of the reference's ``dist_checkpointing/{dict_utils,utils,mapping}.py`` that model code uses inside
``sharded_state_dict`` implementations (prefix rewriting, extraction by type, in-place maps, tensor factories)."""
from __future__ import annotations

import dataclasses
from collections import defaultdict
from itertools import zip_longest
from typing import Any, Callable, Dict, Iterable, List, Optional, Tuple, Union

import torch

from ..ckpt.serialization import LocalNonpersistentObject
from ..ckpt.sharded import ShardedObject, ShardedTensor

ShardedBase = (ShardedTensor, ShardedObject)          # ``isinstance(x, ShardedBase)`` works with the tuple
ShardedStateDict = Dict[str, Any]


# -- dict_utils.py -------------------------------------------------------------------------------------------------------------
def extract_matching_values(x: Union[dict, list], predicate: Callable[[Any], bool], return_lists_as_dicts: bool = False):
    """Split a nested dict / list into (matching, non-matching) structures of the same shape."""
    if isinstance(x, dict):
        yes, no = {}, {}
        for k, v in x.items():
            if isinstance(v, (dict, list)):
                a, b = extract_matching_values(v, predicate, return_lists_as_dicts)
                if a:
                    yes[k] = a
                if b or not a:
                    no[k] = b
            elif predicate(v):
                yes[k] = v
            else:
                no[k] = v
        return yes, no
    if isinstance(x, list):
        yes, no = ({}, {}) if return_lists_as_dicts else ([], [])

        def add(dst, i, val):
            if return_lists_as_dicts:
                dst[i] = val
            else:
                dst.append(val)

        for i, v in enumerate(x):
            if isinstance(v, (dict, list)) and v:
                a, b = extract_matching_values(v, predicate, return_lists_as_dicts)
                if a:
                    add(yes, i, a)
                if b or not a:
                    add(no, i, b)
            else:
                add(yes if predicate(v) else no, i, v)
        return yes, no
    raise ValueError(f"unexpected top-level object type: {type(x)}")


def diff(x1: Any, x2: Any, prefix: Tuple = ()) -> Tuple[list, list, list]:
    """Keys only in the left, only in the right, and with differing values (as key paths)."""
    only_left, only_right, mismatch = [], [], []
    if isinstance(x1, dict) and isinstance(x2, dict):
        only_left = [prefix + (k,) for k in x1.keys() - x2.keys()]
        only_right = [prefix + (k,) for k in x2.keys() - x1.keys()]
        for k in x2.keys() & x1.keys():
            a, b, c = diff(x1[k], x2[k], prefix + (k,))
            only_left += a; only_right += b; mismatch += c
    elif isinstance(x1, (list, tuple)) and isinstance(x2, (list, tuple)):
        if len(x1) != len(x2):
            mismatch.append((prefix + ("length",), len(x1), len(x2)))
        for i, (a1, a2) in enumerate(zip(x1, x2)):
            a, b, c = diff(a1, a2, prefix + (i,))
            only_left += a; only_right += b; mismatch += c
    else:
        if isinstance(x1, torch.Tensor) and isinstance(x2, torch.Tensor):
            same = x1.shape == x2.shape and bool(torch.all(x1.cpu() == x2.cpu()))
        else:
            try:
                same = bool(x1 == x2)
            except RuntimeError:
                same = False
        if not same:
            mismatch.append((prefix, type(x1), type(x2)))
    return only_left, only_right, mismatch


def inspect_types(x: Any, prefix: Tuple = (), indent: int = 4):
    """Print the type structure of a nested state dict (debugging)."""
    pad = " " * indent * len(prefix)
    if isinstance(x, dict):
        for k, v in x.items():
            print(f"{pad}> {k}:" if isinstance(v, (dict, list)) else f"{pad}> {k}: {_describe(v)}")
            if isinstance(v, (dict, list)):
                inspect_types(v, prefix + (k,), indent)
    elif isinstance(x, list):
        for i, v in enumerate(x):
            print(f"{pad}- {i}:" if isinstance(v, (dict, list)) else f"{pad}- {i}: {_describe(v)}")
            if isinstance(v, (dict, list)):
                inspect_types(v, prefix + (i,), indent)
    else:
        print(pad + _describe(x))


def _describe(v) -> str:
    return f"Tensor of shape {tuple(v.shape)}" if isinstance(v, torch.Tensor) else (type(v).__name__ if not isinstance(v, (str, int, float)) else repr(v))


def nested_values(x: Union[dict, list]) -> Iterable[Any]:
    for v in (x.values() if isinstance(x, dict) else x):
        if isinstance(v, (dict, list)):
            yield from nested_values(v)
        else:
            yield v


def nested_items_iter(x: Union[dict, list]):
    """Yield ``(container, key, value)`` for every leaf: the leaf can be replaced through ``container[key] = ...``."""
    for k, v in (x.items() if isinstance(x, dict) else enumerate(x)):
        if isinstance(v, (dict, list)):
            yield from nested_items_iter(v)
        else:
            yield x, k, v


def dict_map(f: Callable, d: dict) -> None:
    for sub, k, v in nested_items_iter(d):
        sub[k] = f(v)


def dict_map_with_key(f: Callable, d: dict) -> None:
    for sub, k, v in nested_items_iter(d):
        sub[k] = f(k, v)


def dict_list_map_inplace(f: Callable, x: Union[dict, list]):
    if isinstance(x, dict):
        for k, v in x.items():
            x[k] = dict_list_map_inplace(f, v)
    elif isinstance(x, list):
        x[:] = (dict_list_map_inplace(f, v) for v in x)
    else:
        return f(x)
    return x


def dict_list_map_outplace(f: Callable, x: Union[dict, list]):
    if isinstance(x, dict):
        return {k: dict_list_map_outplace(f, v) for k, v in x.items()}
    if isinstance(x, list):
        return [dict_list_map_outplace(f, v) for v in x]
    return f(x)


def merge(x1: Union[dict, list], x2: Union[dict, list], key: Tuple = ()):
    """Merge ``x2`` into ``x1`` (in place): dicts recursively, lists element-wise; conflicting leaves are an error."""
    if isinstance(x1, dict) and isinstance(x2, dict):
        for k, v2 in x2.items():
            if k not in x1:
                x1[k] = v2
            else:
                x1[k] = merge(x1[k], v2, key + (k,))
    elif isinstance(x1, list) and isinstance(x2, list):
        if len(x1) != len(x2):
            raise ValueError(f"cannot merge two lists of different lengths ({len(x1)} and {len(x2)}) at {key}")
        for i, v2 in enumerate(x2):
            x1[i] = merge(x1[i], v2, key + (i,))
    else:
        raise ValueError(f"duplicate non-dict and non-list values encountered at {key}: `{x1}` and `{x2}`")
    return x1


def map_reduce(xs: Iterable, key_fn: Callable = lambda x: x, value_fn: Callable = lambda x: x, reduce_fn: Callable = lambda x: x) -> dict:
    groups = defaultdict(list)
    for x in xs:
        groups[key_fn(x)].append(value_fn(x))
    return {k: reduce_fn(v) for k, v in groups.items()}


# -- utils.py -----------------------------------------------------------------------------------------------------------------
def zip_strict(*args):
    """``zip`` that refuses iterables of different lengths."""
    sentinel = object()
    for group in zip_longest(*args, fillvalue=sentinel):
        if any(v is sentinel for v in group):
            raise ValueError("zip_strict: iterables have different lengths")
        yield group


def extract_sharded_tensors(sharded_state_dict):
    return extract_matching_values(sharded_state_dict, lambda v: isinstance(v, ShardedTensor))


def extract_sharded_tensors_and_factories(sharded_state_dict):
    return extract_matching_values(sharded_state_dict, lambda v: isinstance(v, (ShardedTensor, ShardedTensorFactory)))


def extract_sharded_tensors_or_nonpersistent(sharded_state_dict):
    return extract_matching_values(sharded_state_dict,
                                   lambda v: isinstance(v, (ShardedTensor, LocalNonpersistentObject, ShardedTensorFactory)))


def extract_sharded_base(sharded_state_dict):
    return extract_matching_values(sharded_state_dict, lambda v: isinstance(v, ShardedBase))


def extract_nonpersistent(sharded_state_dict):
    return extract_matching_values(sharded_state_dict, lambda v: isinstance(v, LocalNonpersistentObject))


def _rekey(obj, fn: Callable[[str], str]):
    if isinstance(obj, (ShardedTensor, ShardedObject, ShardedTensorFactory)):
        obj.key = fn(obj.key)
    return obj


def add_prefix_for_sharding(sharded_state_dict, prefix: str):
    """Prepend ``prefix`` to the checkpoint key of every sharded entry (in place)."""
    dict_list_map_inplace(lambda t: _rekey(t, lambda k: f"{prefix}{k}"), sharded_state_dict)


def replace_prefix_for_sharding(sharded_state_dict, old_prefix: str, new_prefix: str):
    def fn(k: str) -> str:
        if not k.startswith(old_prefix):
            raise ValueError(f"expected {k} to begin with prefix {old_prefix}")
        return f"{new_prefix}{k[len(old_prefix):]}"

    dict_list_map_inplace(lambda t: _rekey(t, fn), sharded_state_dict)


def apply_prefix_mapping(sharded_state_dict, prefix_map: Dict[str, str]):
    """Rewrite keys by the first matching ``old prefix -> new prefix`` pair."""
    def fn(k: str) -> str:
        for old, new in prefix_map.items():
            if k.startswith(old):
                return f"{new}{k[len(old):]}"
        return k

    dict_list_map_inplace(lambda t: _rekey(t, fn), sharded_state_dict)


# -- mapping.py ---------------------------------------------------------------------------------------------------------------
def is_main_replica(replica_id) -> bool:
    return replica_id == 0 or (isinstance(replica_id, (tuple, list)) and all(r == 0 for r in replica_id))


@dataclasses.dataclass
class ShardedTensorFactory:
    """A tensor whose checkpoint representation is produced by ``build_fn(key, data, replica_id, flattened_range)`` (one or more
    ShardedTensors, e.g. the gate / up halves of a fused SwiGLU weight) and folded back by ``merge_fn`` after loading."""
    key: str
    data: torch.Tensor
    build_fn: Callable
    merge_fn: Callable
    replica_id: Any = 0
    flattened_range: Optional[slice] = None

    def build(self):
        return self.build_fn(self.key, self.data, self.replica_id, self.flattened_range)

    def without_data(self):
        return dataclasses.replace(self, data=None)


def apply_factories(sharded_state_dict):
    """Replace every factory by what it builds (in place)."""
    dict_list_map_inplace(lambda x: x.build() if isinstance(x, ShardedTensorFactory) else x, sharded_state_dict)


def apply_factory_merges(x1, x2, key: Tuple = ()):
    """``x1``: loaded state dict, ``x2``: the factories (same structure): every loaded sub-structure that came from a factory is
    folded back with its ``merge_fn``."""
    if isinstance(x2, ShardedTensorFactory):
        return x2.merge_fn(x1)
    if isinstance(x1, dict) and isinstance(x2, dict):
        for k, v2 in x2.items():
            if k not in x1:
                raise ValueError(f"different dict keys encountered in `apply_factory_merges` ({sorted(x1)} vs {sorted(x2)})")
            x1[k] = apply_factory_merges(x1[k], v2, key + (k,))
    elif isinstance(x1, list) and isinstance(x2, list):
        if len(x1) != len(x2):
            raise ValueError(f"cannot merge two lists with different lengths ({len(x1)} and {len(x2)}) at {key}")
        for i, v2 in enumerate(x2):
            x1[i] = apply_factory_merges(x1[i], v2, key + (i,))
    elif isinstance(x1, list) and isinstance(x2, dict):
        for k, v2 in x2.items():
            x1[int(k)] = apply_factory_merges(x1[int(k)], v2, key + (k,))
    else:
        raise ValueError(f"duplicate non-dict and non-list values encountered at {key}")
    return x1
