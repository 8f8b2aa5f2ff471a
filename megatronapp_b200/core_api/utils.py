"""``megatron.core.utils`` (reference ``megatron/core/utils.py``): the helpers user code imports from there."""
from __future__ import annotations

import logging
import operator
from functools import reduce, wraps
from typing import Any, List, Optional

import torch
import torch.distributed as dist

from ..ckpt.sharded import make_sharded_tensor_for_checkpoint, make_tp_sharded_tensor_for_checkpoint  # noqa: F401
from ..config.transformer_config import init_method_normal, scaled_init_method_normal  # noqa: F401
from ..parallel import state as ps
from ..parallel.state import GlobalMemoryBuffer  # noqa: F401
from ..utils.straggler import StragglerDetector  # noqa: F401

logger = logging.getLogger(__name__)


class ExperimentalNotEnabledError(Exception):
    """Raised by ``@experimental_fn`` / ``@experimental_cls`` objects when experimental features are switched off."""


def _experimental_enabled() -> bool:
    try:
        from ..training.global_vars import get_args

        return bool(getattr(get_args(), "enable_experimental", True))
    except Exception:  # noqa: BLE001  (library use)
        return True


def experimental_fn(introduced_with_version: str):
    def deco(fn):
        @wraps(fn)
        def inner(*a, **k):
            if not _experimental_enabled():
                raise ExperimentalNotEnabledError(f"{fn.__name__} is experimental (since {introduced_with_version})")
            return fn(*a, **k)
        return inner
    return deco


def experimental_cls(introduced_with_version: str):
    def deco(cls):
        init = cls.__init__

        @wraps(init)
        def guarded(self, *a, **k):
            if not _experimental_enabled():
                raise ExperimentalNotEnabledError(f"{cls.__name__} is experimental (since {introduced_with_version})")
            init(self, *a, **k)
        cls.__init__ = guarded
        return cls
    return deco


# -- versions -----------------------------------------------------------------------------------------------------
def _version_tuple(v: str):
    out = []
    for part in v.split("+")[0].split("."):
        digits = "".join(ch for ch in part if ch.isdigit())
        out.append(int(digits) if digits else 0)
    return tuple(out)


def get_torch_version():
    return _version_tuple(torch.__version__)


def is_torch_min_version(version: str, check_equality: bool = True) -> bool:
    have, want = get_torch_version(), _version_tuple(version)
    return have >= want if check_equality else have > want


def get_te_version():
    """Transformer Engine is not part of this framework (its layers are native): there is no version."""
    return None


def is_te_min_version(version: str, check_equality: bool = True) -> bool:
    return False


def get_fa_version():
    return None


def is_fa_min_version(version: str, check_equality: bool = True) -> bool:
    return False


# -- arithmetic -----------------------------------------------------------------------------------------------------
def ensure_divisibility(numerator: int, denominator: int) -> None:
    if numerator % denominator != 0:
        raise AssertionError(f"{numerator} is not divisible by {denominator}")


def divide(numerator: int, denominator: int) -> int:
    ensure_divisibility(numerator, denominator)
    return numerator // denominator


# -- model introspection --------------------------------------------------------------------------------------------------
def deprecate_inference_params(inference_context, inference_params):
    """``inference_params`` was renamed ``inference_context``; accept the old keyword."""
    if inference_context is None and inference_params is not None:
        return inference_params
    return inference_context


def get_attr_wrapped_model(model, attr: str, allow_none: bool = True, return_model_obj: bool = False):
    if isinstance(model, list):
        raise RuntimeError("get_attr_wrapped_model takes one model, not a list of virtual-pipeline chunks")
    m = model
    while not hasattr(m, attr):
        if not hasattr(m, "module"):
            if allow_none:
                return None
            raise RuntimeError(f"couldn't find attribute {attr}")
        m = m.module
    return m if return_model_obj else getattr(m, attr)


def get_model_type(model):
    return get_attr_wrapped_model(model, "model_type")


def get_model_xattn(model):
    try:
        return get_attr_wrapped_model(model, "xattn_needed", allow_none=False)
    except RuntimeError:
        return False


def get_model_config(model):
    return get_attr_wrapped_model(model, "config", allow_none=False)


def is_submodule(module: torch.nn.Module, parent_module: torch.nn.Module, strict: bool = True) -> bool:
    if strict and module is parent_module:
        return False
    return any(m is module for m in parent_module.modules())


# -- viewless tensors (activations handed between pipeline stages must own their storage) ---------------------------------
class WrappedTensor:
    """Holds a tensor that the callee may take and free (``unwrap`` hands it over exactly once)."""

    def __init__(self, tensor: torch.Tensor):
        self._wrapper = [tensor]

    def unwrap(self):
        if not self._wrapper:
            raise RuntimeError("WrappedTensor has already been unwrapped")
        return self._wrapper.pop(0)


def _kernel_make_viewless_tensor(inp, requires_grad):
    out = torch.empty((1,), dtype=inp.dtype, device=inp.device, requires_grad=requires_grad)
    out.data = inp.data
    return out


class MakeViewlessTensor(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, requires_grad):
        return _kernel_make_viewless_tensor(inp, requires_grad)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None


def make_viewless_tensor(inp, requires_grad, keep_graph):
    if inp._base is None:
        return inp
    return MakeViewlessTensor.apply(inp, requires_grad) if keep_graph else _kernel_make_viewless_tensor(inp, requires_grad)


def assert_viewless_tensor(tensor, extra_msg=None):
    if isinstance(tensor, list):
        return [assert_viewless_tensor(t) for t in tensor]
    if isinstance(tensor, torch.Tensor) and tensor._base is not None:
        raise AssertionError(f"a view leaked where a tensor owning its memory is required. {extra_msg or ''}")
    return tensor


def safely_set_viewless_tensor_data(tensor, new_data_tensor):
    assert_viewless_tensor(tensor, extra_msg=f"FYI, tensor._base has shape {None if tensor._base is None else tensor._base.shape}")
    tensor.data = new_data_tensor


# -- logging --------------------------------------------------------------------------------------------------------
def log_single_rank(logger_: logging.Logger, *args: Any, rank: int = 0, **kwargs: Any):
    if not dist.is_initialized() or dist.get_rank() == rank:
        logger_.log(*args, **kwargs)


def log_on_each_pipeline_stage(logger_: logging.Logger, *args: Any, **kwargs: Any):
    if not dist.is_initialized() or (ps.get_data_parallel_rank(with_context_parallel=True) == 0
                                     and ps.get_tensor_model_parallel_rank() == 0):
        logger_.log(*args, **kwargs)


def check_param_hashes_across_dp_replicas(model: List[torch.nn.Module], cross_check: bool = False) -> bool:
    from ..training.training import check_param_hashes_across_dp_replicas as impl

    return impl(model)


# -- DTensor helpers (FSDP2 parameters) ------------------------------------------------------------------------------------
def _is_dtensor(t) -> bool:
    try:
        from torch.distributed.tensor import DTensor
    except ImportError:      # pragma: no cover
        return False
    return isinstance(t, DTensor)


def get_full_tensor_if_necessary(tensor):
    return tensor.full_tensor() if _is_dtensor(tensor) else tensor


def to_local_if_dtensor(tensor):
    return tensor.to_local() if _is_dtensor(tensor) else tensor


def get_data_parallel_group_if_dtensor(tensor, data_parallel_group=None):
    if _is_dtensor(tensor):
        return tensor.device_mesh.get_group()
    return data_parallel_group


# -- weight-gradient helpers ---------------------------------------------------------------------------------------------
def prepare_input_tensors_for_wgrad_compute(grad_output, all_gathered_input):
    """[s, b, h] -> [s*b, h] for both operands of ``dW = dY^T X``."""
    if grad_output.dim() == 3:
        grad_output = grad_output.contiguous().view(-1, grad_output.shape[-1])
        all_gathered_input = all_gathered_input.contiguous().view(-1, all_gathered_input.shape[-1])
    return grad_output, all_gathered_input


def drain_embedding_wgrad_compute(config, embedding_activation_buffer, grad_output_buffer, weight, tp_group=None):
    """Deferred output-layer weight gradients (``defer_embedding_wgrad_compute``): for every stashed (input, grad_output)
    pair gather the sequence-parallel input and accumulate ``dW += dY^T X`` into ``weight.main_grad`` (tcgen05 wgrad)."""
    from .. import ops
    from ..parallel.mappings import gather_from_sequence_parallel_region

    if len(embedding_activation_buffer) != len(grad_output_buffer):
        raise ValueError("activation and gradient buffers differ in length")
    while embedding_activation_buffer:
        x, dy = embedding_activation_buffer.pop(0), grad_output_buffer.pop(0)
        if getattr(config, "sequence_parallel", False) and ps.get_tensor_model_parallel_world_size() > 1:
            x = gather_from_sequence_parallel_region(x, tensor_parallel_output_grad=False)
        dy2, x2 = prepare_input_tensors_for_wgrad_compute(dy, x)
        fuse = hasattr(weight, "main_grad") and getattr(config, "gradient_accumulation_fusion", False)
        dw = ops.wgrad(dy2, x2, weight, fuse)
        if dw is not None:
            weight.grad = dw if weight.grad is None else weight.grad + dw


# -- multi-tensor ops (apex ``multi_tensor_applier`` call shape) --------------------------------------------------------------
def local_multi_tensor_applier(op, noop_flag_buffer, tensor_lists, *args):
    return op(2048 * 32, noop_flag_buffer, tensor_lists, *args)


def local_multi_tensor_l2_norm(chunk_size, noop_flag, tensor_lists, per_tensor, *args):
    from .. import ops

    ts = [t for t in tensor_lists[0] if t.numel()]
    total = ops.l2norm_sq(ts).sqrt() if ts else torch.zeros(1, device=noop_flag.device if torch.is_tensor(noop_flag) else "cpu")
    per = torch.stack([t.float().norm() for t in ts]) if (per_tensor and ts) else None
    return total.reshape(1), per


def local_multi_tensor_scale(chunk_size, noop_flag, tensor_lists, scale):
    for src, dst in zip(tensor_lists[0], tensor_lists[1]):
        dst.copy_(src * scale)


def get_batch_on_this_cp_rank(batch):
    from ..parallel.context_parallel import get_batch_on_this_cp_rank as impl

    return impl(batch)


def get_tensor_model_parallel_group_if_none(tp_group, is_expert: bool = False, check_initialized: bool = True):
    if tp_group is not None:
        return tp_group
    return ps.get_expert_tensor_parallel_group() if is_expert else ps.get_tensor_model_parallel_group()


def numel(shape) -> int:
    return reduce(operator.mul, shape, 1)
