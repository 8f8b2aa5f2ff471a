"""Tensor-parallel / sequence-parallel linear layers and the vocab-parallel embedding.

API parity with reference ``megatron/core/tensor_parallel/layers.py`` (``ColumnParallelLinear``
:675, ``RowParallelLinear`` :1019, ``VocabParallelEmbedding`` :172; both linears take/return
``[s, b, h]`` and return ``(output, bias)``).

Hot path on B200 (when a :class:`~megatronapp_b200.comm.arena.TPFusedComm` is installed and the
token count is a multiple of ``tp*128``):

* column-parallel forward  = ONE kernel ``all-gather(x) -> GEMM``          (``gemm_ag``)
* row-parallel forward     = ONE kernel ``GEMM -> reduce-scatter``          (``gemm_rs``)
* column-parallel dgrad    = ONE kernel ``dY W -> reduce-scatter``          (``gemm_rs``, B MN-major)
* row-parallel dgrad       = ONE kernel ``all-gather(dY) -> dY W``          (``gemm_ag``), whose
  gathered dY view feeds the wgrad GEMM directly
* wgrad                    = tcgen05 GEMM with fp32 ``+=`` epilogue into ``main_grad``

Otherwise the same math runs as tcgen05 GEMM + NCCL collectives (the reference's algorithm,
``layers.py:429-443,461-552,1200-1217``), which is also the in-tree baseline.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import state as ps
from .mappings import (
    _group_ranks,
    gather_from_sequence_parallel_region,
    _gather_along_first_dim,
    _reduce_scatter_along_first_dim,
    copy_to_tensor_model_parallel_region,
    gather_from_tensor_model_parallel_region,
    reduce_from_tensor_model_parallel_region,
    reduce_scatter_to_sequence_parallel_region,
    scatter_to_tensor_model_parallel_region,
)
from .random import get_cuda_rng_tracker
from .. import ops
from ..scan.tracer import scope as _scope
from ..scope import hooks as scope_hooks

_FUSED_COMM = None  # TPFusedComm for the tensor-parallel group (installed by setup_tp_fused_comm)


def setup_tp_fused_comm(max_tokens: int, hidden: int, max_out: int, n_comm_ctas: int = 8):
    """Create the TP symmetric arena and enable the fused AG/RS GEMM kernels."""
    global _FUSED_COMM
    from ..comm.arena import SymmetricArena, TPFusedComm

    tp = ps.get_tensor_model_parallel_world_size()
    if tp < 2 or not torch.cuda.is_available():
        _FUSED_COMM = None
        return None
    max_dim = max(hidden, max_out)
    nbytes = 2 * (max_tokens * max_dim * 2) * 2 + 4 * 1024 * 1024
    nbytes += 2 * (max_tokens * hidden * 2 // tp + 4096)        # K14: two embedding-output slots
    arena = SymmetricArena(ps.get_tensor_model_parallel_group(), nbytes)
    _FUSED_COMM = TPFusedComm(arena, max_tokens, max_dim, max_dim, n_comm_ctas)
    return _FUSED_COMM


def disable_tp_fused_comm():
    global _FUSED_COMM
    _FUSED_COMM = None


def get_tp_fused_comm():
    return _FUSED_COMM


def _fusable(rows_total: int, *tensors) -> bool:
    fc = _FUSED_COMM
    if fc is None:
        return False
    tp = fc.arena.size
    if rows_total % (tp * 128) != 0:
        return False
    for t in tensors:
        if not (t.is_cuda and t.dtype in (torch.bfloat16, torch.float16) and t.shape[-1] % 8 == 0):
            return False
    return True


def set_tensor_model_parallel_attributes(tensor, is_parallel, dim, stride):
    tensor.tensor_model_parallel = is_parallel
    tensor.partition_dim = dim
    tensor.partition_stride = stride


def param_is_not_tensor_parallel_duplicate(param):
    return getattr(param, "tensor_model_parallel", False) or ps.get_tensor_model_parallel_rank() == 0


def _init_weight(weight, init_method, partition_dim, stride=1, is_expert=False):
    set_tensor_model_parallel_attributes(weight, True, partition_dim, stride)
    name = "expert-parallel-rng" if is_expert else "model-parallel-rng"
    tracker = get_cuda_rng_tracker()
    if name in tracker.states_:
        with tracker.fork(name):
            init_method(weight)
    else:
        init_method(weight)


def _fp8_recipe_from_config(config, is_expert: bool = False):
    """``config.fp8`` in {"e4m3", "hybrid"} turns the linear layers' GEMMs into fp8 (reference
    ``fp8_utils.get_fp8_context``); the embedding / output layer stay in 16-bit like TE."""
    fmt = getattr(config, "fp8", None)
    if not fmt:
        return None
    from ..ops.fp8 import Fp8Recipe

    rec = getattr(config, "fp8_recipe", "delayed")
    return Fp8Recipe(fmt="hybrid" if fmt == "hybrid" else "e4m3", recipe=rec if rec in ("delayed", "mxfp8") else "tensorwise",
                     margin=int(getattr(config, "fp8_margin", 0) or 0),
                     amax_history_len=int(getattr(config, "fp8_amax_history_len", 1) or 1),
                     amax_compute_algo=getattr(config, "fp8_amax_compute_algo", "most_recent"),
                     wgrad=bool(getattr(config, "fp8_wgrad", True)))


# ---------------------------------------------------------------------------
# column parallel
# ---------------------------------------------------------------------------

class _ColumnParallelFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, sequence_parallel, fuse_wgrad, allreduce_dgrad, act=0):
        ctx.sp, ctx.fuse_wgrad, ctx.allreduce_dgrad = sequence_parallel, fuse_wgrad, allreduce_dgrad
        ctx.has_bias = bias is not None
        ctx.act = act
        ctx.x_shape = x.shape
        x2 = x.reshape(-1, x.shape[-1])
        tp = ps.get_tensor_model_parallel_world_size()
        ctx.fused = False
        if sequence_parallel and tp > 1:
            rows_total = x2.shape[0] * tp
            # bias + tanh-GELU ride in the epilogue of the fused kernel too; the pre-activation is kept for backward
            pre = torch.empty(rows_total, weight.shape[0], dtype=x2.dtype, device=x2.device) if act else None
            if _fusable(rows_total, x2, weight):
                with _scope("ag_gemm", group=_group_ranks, expect="ag_gemm", data=x2.numel() * x2.element_size() * (tp - 1)):
                    out, _ = _FUSED_COMM.ag_gemm(x2.contiguous(), weight, bias, act=act, out_pre=pre)
                ctx.fused = True
            else:
                total = _gather_along_first_dim(x2)
                out = ops.gemm(total, weight, bias=bias, act=act, out_pre=pre)
            out_shape = (x.shape[0] * tp,) + tuple(x.shape[1:-1]) + (weight.shape[0],)
            if act:
                ctx.save_for_backward(x2, weight, pre)
                return out.view(out_shape)
        else:
            pre = None
            if act:
                # bias + tanh-GELU applied in the GEMM epilogue; the pre-activation is kept for backward
                pre = torch.empty(x2.shape[0], weight.shape[0], dtype=x2.dtype, device=x2.device)
            out = ops.gemm(x2, weight, bias=bias, act=act, out_pre=pre)
            out_shape = tuple(x.shape[:-1]) + (weight.shape[0],)
            if act:
                ctx.save_for_backward(x2, weight, pre)
                return out.view(out_shape)
        ctx.save_for_backward(x2, weight)
        return out.view(out_shape)

    @staticmethod
    def backward(ctx, dy):
        if ctx.act:
            x2, weight, pre = ctx.saved_tensors
        else:
            x2, weight = ctx.saved_tensors
        tp = ps.get_tensor_model_parallel_world_size()
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        if ctx.act:
            dy2 = ops.ext().bias_gelu_bwd(dy2, pre, None)    # d(pre); the bias is already inside ``pre``
        handle = None
        total = x2
        group = ps.get_tensor_model_parallel_group()
        # fused AG-wgrad: the activation re-gather rides inside the wgrad GEMM (its reduction axis is the token axis), so no
        # NCCL all-gather competes with the dgrad->reduce-scatter kernel for NVLink and SMs
        main_grad = getattr(weight, "main_grad", None)
        fused_wgrad = (ctx.sp and tp > 1 and ctx.needs_input_grad[1] and ctx.fused and ctx.fuse_wgrad
                       and main_grad is not None and main_grad.dtype == torch.float32 and _FUSED_COMM is not None
                       and _FUSED_COMM.use_fused_ag_wgrad and _fusable(dy2.shape[0], dy2, x2)
                       and weight.shape[0] >= 256 and x2.is_contiguous())
        if ctx.sp and tp > 1 and ctx.needs_input_grad[1] and not fused_wgrad:
            # re-gather the activation for wgrad, overlapped with the dgrad GEMM (reference :461-471)
            total = torch.empty((x2.shape[0] * tp, x2.shape[1]), dtype=x2.dtype, device=x2.device)
            handle = dist.all_gather_into_tensor(total, x2, group=group, async_op=True)
        dx = None
        ar_handle = rs_handle = None
        if ctx.needs_input_grad[0]:
            if ctx.sp and tp > 1:
                if ctx.fused and _FUSED_COMM.use_fused_rs and _fusable(dy2.shape[0], dy2, weight):
                    with _scope("gemm_rs_dgrad", group=_group_ranks, expect="gemm_rs_dgrad", data=dy2.shape[0] * weight.shape[1] * 2):
                        dx = _FUSED_COMM.gemm_rs(dy2, weight, b_mn=True)
                else:
                    # dgrad reduce-scatter runs on the NCCL stream underneath the wgrad GEMM
                    # (reference layers.py:496-507,548-552)
                    dfull = ops.gemm(dy2, weight, b_mn=True)
                    dx = torch.empty((dfull.shape[0] // tp, dfull.shape[1]), dtype=dfull.dtype, device=dfull.device)
                    if dfull.is_cuda and dist.get_backend(group) != "gloo":
                        rs_handle = dist.reduce_scatter_tensor(dx, dfull, group=group, async_op=True)
                    else:
                        dx = _reduce_scatter_along_first_dim(dfull)
            else:
                dx = ops.gemm(dy2, weight, b_mn=True)
                if ctx.allreduce_dgrad and tp > 1:
                    ar_handle = dist.all_reduce(dx, group=group, async_op=True)
            dx = dx.view(ctx.x_shape)
        dw = db = None
        if ctx.needs_input_grad[1] and fused_wgrad:
            first = (not getattr(weight, "grad_added_to_main_grad", False)) and getattr(weight, "b200_wgrad_store_first", False)
            with _scope("ag_wgrad", group=_group_ranks, expect="ag_wgrad", data=x2.numel() * x2.element_size() * (tp - 1)):
                _FUSED_COMM.ag_wgrad(dy2, x2, main_grad, accumulate=not first)
            weight.grad_added_to_main_grad = True
            if getattr(weight, "b200_wgrad_store_first", False):
                hook = getattr(weight, "grad_ready_hook", None)
                if hook is not None:
                    hook()
        elif ctx.needs_input_grad[1]:
            if handle is not None:
                handle.wait()
            dw = ops.wgrad(dy2, total, weight, ctx.fuse_wgrad)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = ops.bias_grad(dy2).to(dy.dtype)
        if ar_handle is not None:
            ar_handle.wait()
        if rs_handle is not None:
            rs_handle.wait()
        return dx, dw, db, None, None, None, None


def linear_with_grad_accumulation_and_async_allreduce(
    input, weight, bias, gradient_accumulation_fusion, allreduce_dgrad=None, sequence_parallel=False,
    grad_output_buffer=None, wgrad_deferral_limit=0, async_grad_allreduce=None,
):
    if allreduce_dgrad is None:
        allreduce_dgrad = bool(async_grad_allreduce)
    return _ColumnParallelFn.apply(input, weight, bias, sequence_parallel, gradient_accumulation_fusion,
                                   allreduce_dgrad)


def linear_with_frozen_weight(input, weight, bias, gradient_accumulation_fusion=False, allreduce_dgrad=None,
                              sequence_parallel=False, tp_group=None, grad_output_buffer=None,
                              wgrad_deferral_limit=None, async_grad_allreduce=None):
    """Column-parallel linear whose weight does not train (reference ``tensor_parallel/layers.py:295-401``):
    same forward and dgrad (+ its collective) as the trainable path, no wgrad GEMM and no activation re-gather."""
    if weight.requires_grad:
        raise RuntimeError("linear_with_frozen_weight: weight.requires_grad must be False")
    if allreduce_dgrad is None:
        allreduce_dgrad = bool(async_grad_allreduce)
    return _ColumnParallelFn.apply(input, weight, bias, sequence_parallel, False, allreduce_dgrad)


class ColumnParallelLinear(nn.Module):
    """Y = X A^T + b with A split along its output dim: A = [A_1; ...; A_p]."""

    def __init__(self, input_size, output_size, *, config, init_method: Callable, bias=True,
                 gather_output=False, stride=1, keep_master_weight_for_test=False, skip_bias_add=False,
                 skip_weight_param_allocation=False, embedding_activation_buffer=None,
                 grad_output_buffer=None, is_expert=False, tp_comm_buffer_name=None,
                 disable_grad_reduce=False, tp_group=None):
        super().__init__()
        self.input_size, self.output_size = input_size, output_size
        self.gather_output, self.skip_bias_add, self.is_expert = gather_output, skip_bias_add, is_expert
        self.config = config
        self.disable_grad_reduce = disable_grad_reduce
        if is_expert:
            ws = ps.get_expert_tensor_parallel_world_size()
        else:
            ws = ps.get_tensor_model_parallel_world_size()
        self.tp_size = ws
        self.explicit_expert_comm = is_expert and (ws > 1 or config.expert_model_parallel_size > 1)
        if output_size % ws != 0:
            raise ValueError(f"output_size {output_size} not divisible by tp {ws}")
        self.output_size_per_partition = output_size // ws
        dev = "cpu" if (config.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        if not skip_weight_param_allocation:
            self.weight = nn.Parameter(torch.empty(self.output_size_per_partition, input_size,
                                                   dtype=config.params_dtype, device=dev))
            if config.perform_initialization:
                _init_weight(self.weight, init_method, 0, stride, is_expert)
            else:
                set_tensor_model_parallel_attributes(self.weight, True, 0, stride)
            setattr(self.weight, "allreduce", not (is_expert and config.expert_model_parallel_size > 1))
            if config.gradient_accumulation_fusion and not is_expert:
                self.weight.b200_wgrad_store_first = True
        else:
            self.weight = None
        if bias:
            self.bias = nn.Parameter(torch.zeros(self.output_size_per_partition, dtype=config.params_dtype, device=dev))
            set_tensor_model_parallel_attributes(self.bias, True, 0, stride)
            setattr(self.bias, "allreduce", not (is_expert and config.expert_model_parallel_size > 1))
        else:
            self.register_parameter("bias", None)
        self.sequence_parallel = config.sequence_parallel and not self.explicit_expert_comm
        self.allreduce_dgrad = ws > 1 and not self.sequence_parallel and not disable_grad_reduce \
            and not self.explicit_expert_comm
        self.gradient_accumulation_fusion = config.gradient_accumulation_fusion
        self.layer_number = None
        self.fp8_recipe = None if skip_weight_param_allocation else _fp8_recipe_from_config(config, is_expert)
        self._fp8_cache = None

    def _fp8_forward(self, input_, weight):
        """fp8 GEMMs around explicit TP collectives (all-gather / copy in front, nothing behind)."""
        from ..ops.fp8 import _MxWeightCache, _WeightCache, fp8_linear

        if self._fp8_cache is None:
            self._fp8_cache = _MxWeightCache() if self.fp8_recipe.recipe == "mxfp8" else _WeightCache()
        if self.sequence_parallel and self.tp_size > 1:
            x = gather_from_sequence_parallel_region(input_, tensor_parallel_output_grad=True)
        elif self.allreduce_dgrad:
            x = copy_to_tensor_model_parallel_region(input_)
        else:
            x = input_
        bias = self.bias if not self.skip_bias_add else None
        return fp8_linear(x, weight, bias, self.fp8_recipe, self._fp8_cache,
                          self.gradient_accumulation_fusion and weight.requires_grad)

    def can_fuse_gelu(self, input_) -> bool:
        """bias + tanh-GELU can ride in the GEMM epilogue (plain and fused all-gather kernels alike)."""
        return (self.fp8_recipe is None and input_.is_cuda and input_.dtype in (torch.bfloat16, torch.float16)
                and not self.gather_output and self.weight is not None
                and self.weight.shape[0] % 64 == 0 and not self.explicit_expert_comm)

    def forward(self, input_, weight=None, runtime_gather_output=None, fused_gelu: bool = False):
        if weight is None:
            if self.weight is None:
                raise RuntimeError("weight was not allocated and none was passed to forward()")
            weight = self.weight
        else:
            exp = (self.output_size_per_partition, self.input_size)
            if tuple(weight.shape) != exp:
                raise RuntimeError(f"supplied weight has shape {tuple(weight.shape)}, expected {exp}")
        weight = scope_hooks.perturb_weight(weight, self)  # MegaScope "storage" perturbation
        if self.fp8_recipe is not None and not fused_gelu and weight is self.weight and input_.shape[-1] % 16 == 0 \
                and weight.shape[0] % 16 == 0 and not self.explicit_expert_comm:
            out = self._fp8_forward(input_, weight)
            gather = self.gather_output if runtime_gather_output is None else runtime_gather_output
            if gather:
                out = gather_from_tensor_model_parallel_region(out)
            return out, (self.bias if self.skip_bias_add else None)
        bias = self.bias if not self.skip_bias_add else None
        if fused_gelu:
            out = _ColumnParallelFn.apply(input_, weight, self.bias, self.sequence_parallel,
                                          self.gradient_accumulation_fusion and weight.requires_grad,
                                          self.allreduce_dgrad, 1)
            return out, None
        out = _ColumnParallelFn.apply(input_, weight, bias, self.sequence_parallel,
                                      self.gradient_accumulation_fusion and weight.requires_grad,
                                      self.allreduce_dgrad)
        gather = self.gather_output if runtime_gather_output is None else runtime_gather_output
        if gather:
            out = gather_from_tensor_model_parallel_region(out)
        return out, (self.bias if self.skip_bias_add else None)

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        from ..ckpt.sharded import make_sharded_tensors_for_checkpoint

        return make_sharded_tensors_for_checkpoint(self.state_dict(prefix="", keep_vars=True), prefix,
                                                   {"weight": 0, "bias": 0}, sharded_offsets)

    def extra_repr(self):
        return f"in={self.input_size}, out={self.output_size}, tp={self.tp_size}, bias={self.bias is not None}"


# ---------------------------------------------------------------------------
# row parallel
# ---------------------------------------------------------------------------

class _RowParallelFn(torch.autograd.Function):
    """x: [s, b, in/tp] -> reduce over TP of x W^T; SP: output [s/tp, b, out]."""

    @staticmethod
    def forward(ctx, x, weight, sequence_parallel, fuse_wgrad, do_reduce):
        ctx.sp, ctx.fuse_wgrad = sequence_parallel, fuse_wgrad
        ctx.x_shape = x.shape
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        tp = ps.get_tensor_model_parallel_world_size()
        ctx.fused = False
        if sequence_parallel and tp > 1 and do_reduce:
            if _fusable(x2.shape[0], x2, weight) and _FUSED_COMM.use_fused_rs:
                with _scope("gemm_rs", group=_group_ranks, expect="gemm_rs", data=x2.shape[0] * weight.shape[0] * 2):
                    out = _FUSED_COMM.gemm_rs(x2, weight)
                ctx.fused = True
            elif _fusable(x2.shape[0], x2, weight):
                out = _reduce_scatter_along_first_dim(ops.gemm(x2, weight))
                ctx.fused = True   # backward still uses the fused AG->GEMM dgrad
            else:
                out = _reduce_scatter_along_first_dim(ops.gemm(x2, weight))
            out_shape = (x.shape[0] // tp,) + tuple(x.shape[1:-1]) + (weight.shape[0],)
        else:
            out = ops.gemm(x2, weight)
            if tp > 1 and do_reduce:
                with _scope("_reduce", group=_group_ranks, expect="_reduce", data=out.numel() * out.element_size()):
                    dist.all_reduce(out, group=ps.get_tensor_model_parallel_group())
            out_shape = tuple(x.shape[:-1]) + (weight.shape[0],)
        ctx.do_reduce = do_reduce
        ctx.save_for_backward(x2, weight)
        return out.view(out_shape)

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        tp = ps.get_tensor_model_parallel_world_size()
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = None
        if ctx.sp and tp > 1 and ctx.do_reduce:
            if ctx.fused and _fusable(dy2.shape[0] * tp, dy2, weight):
                with _scope("ag_gemm_dgrad", group=_group_ranks, expect="ag_gemm_dgrad", data=dy2.numel() * 2 * (tp - 1)):
                    dx, dy_full = _FUSED_COMM.ag_gemm(dy2, weight, None, b_mn=True)
            else:
                dy_full = _gather_along_first_dim(dy2)
                dx = ops.gemm(dy_full, weight, b_mn=True) if ctx.needs_input_grad[0] else None
        else:
            dy_full = dy2
            dx = ops.gemm(dy_full, weight, b_mn=True) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            dw = ops.wgrad(dy_full, x2, weight, ctx.fuse_wgrad)
        if dx is not None:
            dx = dx.view(ctx.x_shape)
        return dx, dw, None, None, None


class RowParallelLinear(nn.Module):
    """Y = X A^T + b with A split along its input dim and X split along its last dim."""

    def __init__(self, input_size, output_size, *, config, init_method: Callable, bias: bool,
                 input_is_parallel: bool, skip_bias_add: bool, stride=1, keep_master_weight_for_test=False,
                 is_expert=False, tp_comm_buffer_name=None, tp_group=None):
        super().__init__()
        self.input_size, self.output_size = input_size, output_size
        self.input_is_parallel, self.skip_bias_add, self.is_expert = input_is_parallel, skip_bias_add, is_expert
        self.config = config
        ws = ps.get_expert_tensor_parallel_world_size() if is_expert else ps.get_tensor_model_parallel_world_size()
        self.tp_size = ws
        self.explicit_expert_comm = is_expert and (ws > 1 or config.expert_model_parallel_size > 1)
        if input_size % ws != 0:
            raise ValueError(f"input_size {input_size} not divisible by tp {ws}")
        self.input_size_per_partition = input_size // ws
        self.sequence_parallel = config.sequence_parallel
        if self.sequence_parallel and not input_is_parallel:
            raise RuntimeError("to enable sequence_parallel, input_is_parallel must be True")
        dev = "cpu" if (config.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        self.weight = nn.Parameter(torch.empty(output_size, self.input_size_per_partition,
                                               dtype=config.params_dtype, device=dev))
        if config.perform_initialization:
            _init_weight(self.weight, init_method, 1, stride, is_expert)
        else:
            set_tensor_model_parallel_attributes(self.weight, True, 1, stride)
        setattr(self.weight, "allreduce", not (is_expert and config.expert_model_parallel_size > 1))
        if config.gradient_accumulation_fusion and not is_expert:
            self.weight.b200_wgrad_store_first = True
        if bias:
            self.bias = nn.Parameter(torch.zeros(output_size, dtype=config.params_dtype, device=dev))
            setattr(self.bias, "allreduce", not (is_expert and config.expert_model_parallel_size > 1))
            setattr(self.bias, "sequence_parallel", self.sequence_parallel)
        else:
            self.register_parameter("bias", None)
        self.gradient_accumulation_fusion = config.gradient_accumulation_fusion
        self.fp8_recipe = _fp8_recipe_from_config(config, is_expert)
        self._fp8_cache = None

    def forward(self, input_):
        if not self.input_is_parallel:
            input_ = scatter_to_tensor_model_parallel_region(input_)
        do_reduce = not self.explicit_expert_comm
        if self.fp8_recipe is not None and input_.shape[-1] % 16 == 0 and self.weight.shape[0] % 16 == 0 \
                and not self.explicit_expert_comm:
            from ..ops.fp8 import _MxWeightCache, _WeightCache, fp8_linear

            if self._fp8_cache is None:
                self._fp8_cache = _MxWeightCache() if self.fp8_recipe.recipe == "mxfp8" else _WeightCache()
            out = fp8_linear(input_, self.weight, None, self.fp8_recipe, self._fp8_cache,
                             self.gradient_accumulation_fusion and self.weight.requires_grad)
            if self.tp_size > 1:
                out = reduce_scatter_to_sequence_parallel_region(out) if self.sequence_parallel \
                    else reduce_from_tensor_model_parallel_region(out)
            if not self.skip_bias_add:
                return (out + self.bias if self.bias is not None else out), None
            return out, self.bias
        out = _RowParallelFn.apply(input_, self.weight, self.sequence_parallel and do_reduce,
                                   self.gradient_accumulation_fusion and self.weight.requires_grad, do_reduce)
        if not self.skip_bias_add:
            out = out + self.bias if self.bias is not None else out
            return out, None
        return out, self.bias

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        from ..ckpt.sharded import make_sharded_tensors_for_checkpoint

        return make_sharded_tensors_for_checkpoint(self.state_dict(prefix="", keep_vars=True), prefix,
                                                   {"weight": 1}, sharded_offsets)

    def extra_repr(self):
        return f"in={self.input_size}, out={self.output_size}, tp={self.tp_size}, bias={self.bias is not None}"


# ---------------------------------------------------------------------------
# vocab-parallel embedding
# ---------------------------------------------------------------------------

class VocabUtility:
    @staticmethod
    def vocab_range_from_per_partition_vocab_size(per_partition, rank, world_size):
        f = rank * per_partition
        return f, f + per_partition

    @staticmethod
    def vocab_range_from_global_vocab_size(global_vocab_size, rank, world_size):
        if global_vocab_size % world_size != 0:
            raise ValueError("vocab size must be divisible by tp size")
        return VocabUtility.vocab_range_from_per_partition_vocab_size(global_vocab_size // world_size, rank, world_size)


class _FusedEmbedSP(torch.autograd.Function):
    """tokens [b, s] -> embeddings [s / tp, b, h] (sequence parallel).  Backward is the adjoint of "lookup then
    reduce-scatter": all-gather dy along s, then the dense embedding gradient of the locally owned rows."""

    @staticmethod
    def forward(ctx, tokens, weight, vocab_start, deterministic):
        ctx.save_for_backward(tokens)
        ctx.vocab_start, ctx.v_local, ctx.deterministic = vocab_start, weight.shape[0], deterministic
        with _scope("embed_scatter", group=_group_ranks, expect="embed_scatter",
                    data=tokens.numel() * weight.shape[1] * 2 // ps.get_tensor_model_parallel_world_size()):
            return _FUSED_COMM.embed(tokens, weight.detach(), vocab_start)

    @staticmethod
    def backward(ctx, dy):
        (tokens,) = ctx.saved_tensors
        dy_full = _gather_along_first_dim(dy.contiguous())                      # [s, b, h]
        local = tokens - ctx.vocab_start
        mask = (local < 0) | (local >= ctx.v_local)
        g = dy_full.transpose(0, 1).masked_fill(mask.unsqueeze(-1), 0.0)        # [b, s, h], rows of foreign tokens zeroed
        dw = torch.ops.aten.embedding_dense_backward(g, local.masked_fill(mask, 0), ctx.v_local, -1, False)
        return None, dw, None, None


class VocabParallelEmbedding(nn.Module):
    def __init__(self, num_embeddings, embedding_dim, *, init_method: Callable,
                 reduce_scatter_embeddings: bool = False, config, tp_group=None):
        super().__init__()
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.reduce_scatter_embeddings = reduce_scatter_embeddings
        self.tp_size = ps.get_tensor_model_parallel_world_size()
        self.vocab_start_index, self.vocab_end_index = VocabUtility.vocab_range_from_global_vocab_size(
            num_embeddings, ps.get_tensor_model_parallel_rank(), self.tp_size)
        self.num_embeddings_per_partition = self.vocab_end_index - self.vocab_start_index
        self.deterministic_mode = config.deterministic_mode
        dev = "cpu" if (config.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        self.weight = nn.Parameter(torch.empty(self.num_embeddings_per_partition, embedding_dim,
                                               dtype=config.params_dtype, device=dev))
        if config.perform_initialization:
            _init_weight(self.weight, init_method, 0, 1)
        else:
            set_tensor_model_parallel_attributes(self.weight, True, 0, 1)

    def forward(self, input_):
        fc = _FUSED_COMM
        if (self.reduce_scatter_embeddings and fc is not None and fc.use_fused_embed and input_.is_cuda and input_.dim() == 2
                and self.weight.element_size() == 2 and input_.shape[1] % self.tp_size == 0
                and fc.embed_fits(input_.numel(), self.embedding_dim)):
            # K14: the owner of a token's row writes it straight into the sequence-parallel output of the rank that
            # owns the position (no masked lookup, no reduce-scatter)
            return _FusedEmbedSP.apply(input_.contiguous(), self.weight, self.vocab_start_index, self.deterministic_mode)
        if self.tp_size > 1:
            mask = (input_ < self.vocab_start_index) | (input_ >= self.vocab_end_index)
            local = (input_ - self.vocab_start_index).masked_fill(mask, 0)
        else:
            local = input_
        out = F.embedding(local, self.weight)
        if self.tp_size > 1:
            out = out.masked_fill(mask.unsqueeze(-1), 0.0)
        if self.reduce_scatter_embeddings:
            # [b, s, h] -> [s, b, h] then reduce-scatter along s (sequence parallel)
            out = out.transpose(0, 1).contiguous()
            return reduce_scatter_to_sequence_parallel_region(out)
        return reduce_from_tensor_model_parallel_region(out)

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        from ..ckpt.sharded import make_sharded_tensors_for_checkpoint

        return make_sharded_tensors_for_checkpoint(self.state_dict(prefix="", keep_vars=True), prefix,
                                                   {"weight": 0}, sharded_offsets, allow_shape_mismatch=("weight",))


# -- attribute helpers / autograd-function names of the reference's ``tensor_parallel/layers.py`` ----------------------------------
_MODEL_PARALLEL_ATTRIBUTE_DEFAULTS = {"tensor_model_parallel": False, "partition_dim": -1, "partition_stride": 1}


def set_defaults_if_not_set_tensor_model_parallel_attributes(tensor):
    for k, v in _MODEL_PARALLEL_ATTRIBUTE_DEFAULTS.items():
        if not hasattr(tensor, k):
            setattr(tensor, k, v)


def copy_tensor_model_parallel_attributes(destination_tensor, source_tensor):
    for k in _MODEL_PARALLEL_ATTRIBUTE_DEFAULTS:
        if hasattr(source_tensor, k):
            setattr(destination_tensor, k, getattr(source_tensor, k))


LinearWithGradAccumulationAndAsyncCommunication = _ColumnParallelFn


class LinearWithFrozenWeight:
    """Name kept for ``.apply`` callers: see ``linear_with_frozen_weight``."""

    @staticmethod
    def apply(input, weight, bias, allreduce_dgrad, tp_group=None):
        return linear_with_frozen_weight(input, weight, bias, allreduce_dgrad=allreduce_dgrad)
