"""``ModuleSpec`` / ``build_module``: config-as-code composition of layers.

Parity: reference ``megatron/core/transformer/spec_utils.py:9-106``.
"""
from __future__ import annotations

import importlib
import types
from dataclasses import dataclass, field
from typing import Tuple, Union


@dataclass
class ModuleSpec:
    module: Union[Tuple, type]
    params: dict = field(default_factory=lambda: {})
    submodules: object = None


def import_module(module_path: Tuple[str, str]):
    base_path, name = module_path
    try:
        module = importlib.import_module(base_path)
    except ImportError as e:
        raise ImportError(f"couldn't import module {base_path}: {e}") from e
    return vars(module)[name]


def get_module(spec_or_module, **additional_kwargs):
    if isinstance(spec_or_module, (type, types.FunctionType)):
        return spec_or_module
    if isinstance(spec_or_module.module, (type, types.FunctionType)):
        return spec_or_module.module
    return import_module(spec_or_module.module)


def build_module(spec_or_module, *args, **kwargs):
    if isinstance(spec_or_module, types.FunctionType):
        return spec_or_module
    if isinstance(spec_or_module, ModuleSpec) and isinstance(spec_or_module.module, types.FunctionType):
        return spec_or_module.module
    if isinstance(spec_or_module, type):
        module = spec_or_module
    elif hasattr(spec_or_module, "module") and isinstance(spec_or_module.module, type):
        module = spec_or_module.module
    else:
        module = import_module(spec_or_module.module)
    if isinstance(module, types.FunctionType):
        return module
    if hasattr(spec_or_module, "submodules") and spec_or_module.submodules is not None:
        kwargs["submodules"] = spec_or_module.submodules
    try:
        params = spec_or_module.params if hasattr(spec_or_module, "params") else {}
        return module(*args, **params, **kwargs)
    except Exception as e:
        import sys

        raise type(e)(f"{e} when instantiating {module.__name__} with args {args} and kwargs {kwargs}"
                      ).with_traceback(sys.exc_info()[2])
