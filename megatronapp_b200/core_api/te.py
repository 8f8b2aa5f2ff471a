"""``megatron.core.extensions.transformer_engine``: the Transformer-Engine layer names that reference layer specs use, bound to
this framework's native modules (there is no TE here -- the tcgen05 / fp8 / fused-collective kernels live behind the same
classes as the "local" spec).  A spec written against TE therefore builds the native layers unchanged."""
from __future__ import annotations

from contextlib import nullcontext
from typing import Optional

import torch

from ..parallel.random import CudaRNGStatesTracker as TECudaRNGStatesTracker  # noqa: F401
from ..parallel.random import checkpoint as _checkpoint
from ..parallel.tp_layers import ColumnParallelLinear, RowParallelLinear
from ..transformer.attention import DotProductAttention
from ..transformer.norms import FusedLayerNorm

HAVE_TE = False          # what ``is_te_min_version`` callers branch on


class TENorm:
    """``TENorm(config, hidden_size, eps)`` -> LayerNorm / RMSNorm according to ``config.normalization``."""

    def __new__(cls, config, hidden_size: int, eps: float = 1e-5):
        return FusedLayerNorm(config, hidden_size, eps)


class TEColumnParallelLinear(ColumnParallelLinear):
    pass


class TERowParallelLinear(RowParallelLinear):
    pass


class TELinear(torch.nn.Module):
    """``parallel_mode`` in {"column", "row", None / "duplicated"}: the matching parallel linear (a replicated linear is a column
    parallel one over a group of size 1)."""

    def __new__(cls, input_size, output_size, *, parallel_mode: Optional[str], config, init_method, bias=True, skip_bias_add=False,
                skip_weight_param_allocation=False, tp_comm_buffer_name=None, is_expert=False, **kw):
        if parallel_mode == "row":
            return RowParallelLinear(input_size, output_size, config=config, init_method=init_method, bias=bias, input_is_parallel=True,
                                     skip_bias_add=skip_bias_add, is_expert=is_expert)
        return ColumnParallelLinear(input_size, output_size, config=config, init_method=init_method, bias=bias, gather_output=False,
                                    skip_bias_add=skip_bias_add, is_expert=is_expert,
                                    skip_weight_param_allocation=skip_weight_param_allocation)


class TELayerNormColumnParallelLinear(torch.nn.Module):
    """LayerNorm followed by a column-parallel linear in one module (TE fuses the two; here the norm kernel feeds the fused
    all-gather -> GEMM).  Parameter names follow TE (``layer_norm_weight`` / ``layer_norm_bias`` / ``weight`` / ``bias``)."""

    def __init__(self, input_size, output_size, *, config, init_method, gather_output=False, bias=True, skip_bias_add=False,
                 is_expert=False, skip_weight_param_allocation=False, tp_comm_buffer_name=None, **kw):
        super().__init__()
        self.norm = FusedLayerNorm(config, input_size, config.layernorm_epsilon)
        self.linear = ColumnParallelLinear(input_size, output_size, config=config, init_method=init_method, bias=bias,
                                           gather_output=gather_output, skip_bias_add=skip_bias_add, is_expert=is_expert)
        self.config = config

    @property
    def weight(self):
        return self.linear.weight

    @property
    def bias(self):
        return self.linear.bias

    @property
    def layer_norm_weight(self):
        return self.norm.weight

    @property
    def layer_norm_bias(self):
        return getattr(self.norm, "bias", None)

    def forward(self, x):
        return self.linear(self.norm(x))

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        from ..ckpt.sharded import make_sharded_tensors_for_checkpoint

        sd = {"layer_norm_weight": self.norm.weight, "weight": self.linear.weight}
        if getattr(self.norm, "bias", None) is not None:
            sd["layer_norm_bias"] = self.norm.bias
        if self.linear.bias is not None:
            sd["bias"] = self.linear.bias
        return make_sharded_tensors_for_checkpoint(sd, prefix, {"weight": 0, "bias": 0}, sharded_offsets)


class TEDotProductAttention(DotProductAttention):
    """The flash-attention kernel (``csrc/attention_sm100.cu``) behind TE's attention class name."""


def TEGroupedLinear(*a, **k):
    raise NotImplementedError("grouped expert linears are part of megatronapp_b200.moe.experts.GroupedMLP (one grouped tcgen05 "
                              "GEMM per matrix); use GroupedMLP / TEGroupedMLP in the layer spec")


TEColumnParallelGroupedLinear = TERowParallelGroupedLinear = TEGroupedLinear


class TEDelayedScaling:
    """fp8 delayed-scaling recipe holder (reference ``TEDelayedScaling``) -> ``ops.fp8.Fp8Recipe``."""

    def __new__(cls, config, fp8_format, override_linear_precision: tuple = (False, False, False)):
        from ..ops.fp8 import Fp8Recipe

        return Fp8Recipe(fmt="hybrid" if str(fp8_format).lower().endswith("hybrid") else "e4m3", recipe="delayed",
                         margin=getattr(config, "fp8_margin", 0), amax_history_len=getattr(config, "fp8_amax_history_len", 1),
                         amax_compute_algo=getattr(config, "fp8_amax_compute_algo", "most_recent"), wgrad=not override_linear_precision[2])


def te_checkpoint(forward_func, distribute_saved_activations, get_rng_state_tracker, tp_group, *args, **kwargs):
    return _checkpoint(lambda *a: forward_func(*a, **kwargs), distribute_saved_activations, *args)


def get_cpu_offload_context(enabled, num_layers, model_layers, activation_offloading, weight_offloading):
    """Activation offloading to the host is not implemented: returns a null context and an identity sync function."""
    if enabled:
        raise NotImplementedError("cpu_offloading of activations is not implemented")
    return nullcontext(), (lambda t: t)


def fused_apply_rotary_pos_emb(t, freqs, transpose_output_memory: bool = False, interleaved: bool = False):
    """``t [s, b, h, d]``, ``freqs [s, 1, 1, rot]`` holding the angles twice (Megatron layout) -> the RoPE kernel's
    ``[s, rot]`` table of ``cos | sin`` over the ``rot / 2`` distinct angles."""
    from .. import ops

    ang = freqs.float().reshape(freqs.shape[0], -1)
    ang = ang[:, : ang.shape[-1] // 2]
    return ops.rope(t, torch.cat((ang.cos(), ang.sin()), dim=-1).contiguous(), interleaved)
