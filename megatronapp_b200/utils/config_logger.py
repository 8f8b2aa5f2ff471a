"""Dump module / optimizer configs to disk for post-mortem diffing (reference ``core/config_logger.py``,
``--config-logger-dir``)."""
import dataclasses
import json
import os

import torch


def has_config_logger_enabled(config) -> bool:
    return bool(getattr(config, "config_logger_dir", ""))


_COUNT = {}


def _jsonable(o):
    if dataclasses.is_dataclass(o) and not isinstance(o, type):
        return {f.name: _jsonable(getattr(o, f.name)) for f in dataclasses.fields(o)}
    if isinstance(o, dict):
        return {str(k): _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(x) for x in o]
    if isinstance(o, (int, float, str, bool)) or o is None:
        return o
    if isinstance(o, torch.dtype):
        return str(o)
    if isinstance(o, torch.nn.Module):
        return {"module": o.__class__.__name__, "repr": repr(o)[:2000]}
    if callable(o):
        return getattr(o, "__name__", repr(o))
    return repr(o)


def log_config_to_disk(config, dict_data, prefix: str = "", rank_str: str = ""):
    path = getattr(config, "config_logger_dir", "")
    if not path:
        return None
    os.makedirs(path, exist_ok=True)
    if not rank_str:
        import torch.distributed as dist

        rank_str = str(dist.get_rank()) if dist.is_initialized() else "0"
    n = _COUNT.get(prefix, 0)
    _COUNT[prefix] = n + 1
    fn = os.path.join(path, f"{prefix}.rank_{rank_str}.iter{n}.json")
    with open(fn, "w") as f:
        json.dump(_jsonable(dict_data), f, indent=1)
    return fn


# -- helper names of the reference's ``core/config_logger.py`` --------------------------------------------------------------------------
def get_config_logger_path(config):
    return getattr(config, "config_logger_dir", "")


def get_path_count(path: str) -> int:
    """How many times ``path`` has been handed out so far (the counter that numbers repeated dumps of one module)."""
    n = _COUNT.get(path, 0)
    _COUNT[path] = n + 1
    return n


def get_path_with_count(path: str) -> str:
    """``<path>.iter<n>`` with ``n`` counting the calls for this path."""
    return f"{path}.iter{get_path_count(path)}"


class JSONEncoderWithMcoreTypes(json.JSONEncoder):
    """``json.dumps(..., cls=JSONEncoderWithMcoreTypes)``: dtypes, dataclasses, modules, callables and enums become strings / dicts."""

    def default(self, o):
        return _jsonable(o)
