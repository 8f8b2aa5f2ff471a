"""Base module classes (reference ``transformer/module.py``: ``MegatronModule`` :25, ``Float16Module`` :137)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..parallel import state as ps

_FLOAT_TYPES = (torch.float32,)
_HALF_TYPES = (torch.float16, torch.bfloat16)


def param_is_not_shared(param):
    return not getattr(param, "shared", False)


def sharded_state_dict_of(module, prefix: str, sharded_offsets=(), metadata=None) -> dict:
    """``module``'s sharded state dict whether or not it implements the protocol itself: containers
    (``ModuleList`` / ``Sequential`` / ``ModuleDict``) are walked so that tensor-parallel leaves below them keep
    their TP axes; plain torch modules contribute replicated tensors."""
    import torch.nn as nn

    from ..ckpt.sharded import make_sharded_tensors_for_checkpoint

    if hasattr(module, "sharded_state_dict"):
        return module.sharded_state_dict(prefix, sharded_offsets, metadata)
    if isinstance(module, (nn.ModuleList, nn.Sequential, nn.ModuleDict)):
        out = {}
        for name, child in module.named_children():
            out.update(sharded_state_dict_of(child, f"{prefix}{name}.", sharded_offsets, metadata))
        return out
    amap = {n: p.partition_dim for n, p in module.named_parameters(recurse=False)
            if getattr(p, "tensor_model_parallel", False) and getattr(p, "partition_dim", -1) >= 0}
    out = make_sharded_tensors_for_checkpoint(
        {**{k: v for k, v in module._parameters.items() if v is not None},
         **{k: v for k, v in module._buffers.items() if v is not None and k not in module._non_persistent_buffers_set}},
        prefix, amap, sharded_offsets)
    for name, child in module.named_children():
        out.update(sharded_state_dict_of(child, f"{prefix}{name}.", sharded_offsets, metadata))
    return out


class MegatronModule(nn.Module):
    def __init__(self, config=None):
        super().__init__()
        self.config = config

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        return self.state_dict(prefix=prefix, keep_vars=keep_vars)

    def log_constructor_args(self, ctor_locals: dict):
        """``--config-logger-dir``: dump the constructor arguments (and through them the full config) of a model
        to disk, one JSON per rank (reference: ``log_config_to_disk(config, locals(), prefix=type(self).__name__)``)."""
        from ..utils.config_logger import has_config_logger_enabled, log_config_to_disk

        if has_config_logger_enabled(self.config):
            data = {k: v for k, v in ctor_locals.items() if k not in ("self", "__class__")}
            log_config_to_disk(self.config, data, prefix=type(self).__name__)

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        from ..ckpt.sharded import make_sharded_tensors_for_checkpoint

        sd = {}
        own = {k: v for k, v in self._parameters.items() if v is not None}
        own.update({k: v for k, v in self._buffers.items() if v is not None and k not in self._non_persistent_buffers_set})
        if own:
            sd.update(make_sharded_tensors_for_checkpoint(own, prefix, {}, sharded_offsets))
        for name, child in self.named_children():
            sd.update(sharded_state_dict_of(child, f"{prefix}{name}.", sharded_offsets, metadata))
        return sd

    def set_is_first_microbatch(self):
        for m in self.modules():
            if hasattr(m, "is_first_microbatch"):
                m.is_first_microbatch = True

    def set_input_tensor(self, input_tensor):
        pass


def _convert(val, fn):
    if isinstance(val, (tuple, list)):
        r = [_convert(v, fn) for v in val]
        return tuple(r) if isinstance(val, tuple) else r
    return fn(val)


class Float16Module(MegatronModule):
    """Casts fp32 inputs to the 16-bit compute dtype on the first stage and 16-bit outputs back
    to fp32 on the last stage."""

    def __init__(self, config, module: nn.Module):
        super().__init__(config)
        if config.fp16:
            self.add_module("module", module.half())
            self._dtype = torch.float16
        elif config.bf16:
            self.add_module("module", module.bfloat16())
            self._dtype = torch.bfloat16
        else:
            raise ValueError("Float16Module needs fp16 or bf16")

    def set_input_tensor(self, input_tensor):
        return self.module.set_input_tensor(input_tensor)

    def forward(self, *inputs, **kwargs):
        if ps.is_pipeline_first_stage():
            to16 = lambda v: v.to(self._dtype) if torch.is_tensor(v) and v.dtype in _FLOAT_TYPES else v  # noqa: E731
            inputs = _convert(inputs, to16)
        out = self.module(*inputs, **kwargs)
        if ps.is_pipeline_last_stage():
            to32 = lambda v: v.float() if torch.is_tensor(v) and v.dtype in _HALF_TYPES else v  # noqa: E731
            out = _convert(out, to32)
        return out

    def state_dict(self, *a, **kw):
        return self.module.state_dict(*a, **kw)

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        return self.module.state_dict_for_save_checkpoint(prefix=prefix, keep_vars=keep_vars)

    def sharded_state_dict(self, prefix="", *args, **kwargs):
        return self.module.sharded_state_dict(prefix, *args, **kwargs)

    def load_state_dict(self, state_dict, strict=True):
        return self.module.load_state_dict(state_dict, strict=strict)


def conversion_helper(val, conversion):
    """Apply ``conversion`` to every element of a (nested) tuple / list."""
    return _convert(val, conversion)


def fp32_to_float16(val, float16_convertor):
    return _convert(val, lambda v: float16_convertor(v) if (torch.is_tensor(v) and v.dtype == torch.float32) else v)


def float16_to_fp32(val):
    return _convert(val, lambda v: v.float() if (torch.is_tensor(v) and v.dtype in (torch.bfloat16, torch.float16)) else v)
