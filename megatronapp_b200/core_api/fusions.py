"""``megatron.core.fusions.*`` call surfaces on the sm_100a elementwise kernels (``csrc/elementwise.cu``): bias + GELU,
bias + SwiGLU / GEGLU, bias + dropout + residual, scaled (masked) softmax, vocabulary-parallel cross entropy."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.nn.functional as F

from .. import ops
from .enums import AttnMaskType


# -- fused_bias_gelu ---------------------------------------------------------------------------------------------------------
def bias_gelu(bias, y):
    return ops.bias_gelu(y, bias)


def bias_gelu_back(g, bias, y):
    x = (y + bias).float()
    t = torch.tanh(0.79788456 * x * (1 + 0.044715 * x * x))
    ff = 0.5 * x * ((1 - t * t) * (0.79788456 + 0.1070322243 * x * x)) + 0.5 * (1 + t)
    return (ff * g.float()).to(g.dtype)


class GeLUFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, bias):
        ctx.save_for_backward(input, bias)
        return bias_gelu(bias, input)

    @staticmethod
    def backward(ctx, grad_output):
        input, bias = ctx.saved_tensors
        tmp = bias_gelu_back(grad_output, bias, input)
        return tmp, tmp.reshape(-1, tmp.shape[-1]).sum(0)

    @classmethod
    def apply(cls, input, bias):            # the kernel-backed op already has an autograd formula; keep ``.apply`` callable
        return ops.bias_gelu(input, bias)


bias_gelu_impl = GeLUFunction.apply


# -- fused_bias_swiglu / geglu -----------------------------------------------------------------------------------------------
def swiglu(y):
    return ops.swiglu(y, None)


def bias_swiglu(y, bias):
    return ops.swiglu(y, bias)


def _glu_back(g, y, act, dact):
    a, b = torch.chunk(y.float(), 2, dim=-1)
    g = g.float()
    return torch.cat((g * dact(a) * b, g * act(a)), dim=-1)


def swiglu_back(g, y):
    sig = torch.sigmoid
    return _glu_back(g, y, F.silu, lambda a: sig(a) * (1 + a * (1 - sig(a)))).to(y.dtype)


def bias_swiglu_back(g, y, bias):
    return swiglu_back(g, y + bias)


class SwiGLUFunction(torch.autograd.Function):
    @classmethod
    def apply(cls, input, fp8_input_store=False, *unused):
        return ops.swiglu(input, None)


class BiasSwiGLUFunction(torch.autograd.Function):
    @classmethod
    def apply(cls, input, bias, fp8_input_store=False, *unused):
        return ops.swiglu(input, bias)


def bias_swiglu_impl(input, bias, fp8_input_store=False, cpu_offload_input=False):
    shape = input.shape
    out = ops.swiglu(input.reshape(-1, shape[-1]), bias)
    return out.view(*shape[:-1], out.shape[-1])


def weighted_bias_swiglu_impl(input, bias, weights, fp8_input_store=False):
    """SwiGLU scaled per token by ``weights [tokens, 1]`` (the router probability of an MoE expert input)."""
    return (bias_swiglu_impl(input, bias).float() * weights.float()).to(input.dtype)


def geglu(y):
    a, b = torch.chunk(y, 2, dim=-1)
    return F.gelu(a, approximate="tanh") * b


def bias_geglu(bias, y):
    return geglu(y + bias)


def bias_geglu_impl(input, bias):
    shape = input.shape
    out = bias_geglu(bias, input.reshape(-1, shape[-1])) if bias is not None else geglu(input.reshape(-1, shape[-1]))
    return out.view(*shape[:-1], out.shape[-1])


def quick_gelu(y):
    return y * torch.sigmoid(1.702 * y)


def quick_geglu(y, linear_offset: float = 0.0):
    a, b = torch.chunk(y, 2, dim=-1)
    return quick_gelu(a) * (b + linear_offset)


def bias_quick_geglu_impl(input, bias, linear_offset: float = 0.0):
    return quick_geglu(input if bias is None else input + bias, linear_offset)


# -- fused_bias_dropout ----------------------------------------------------------------------------------------------------
def _bda(x_with_bias: Tuple[torch.Tensor, Optional[torch.Tensor]], residual, prob: float, training: bool):
    x, bias = x_with_bias
    return ops.bias_dropout_add(x, bias, residual, prob, training)


def bias_dropout_add_unfused(training: bool) -> Callable:
    return lambda x_with_bias, residual, prob: _bda(x_with_bias, residual, prob, training)


def bias_dropout_add_fused_train(x_with_bias, residual, prob):
    return _bda(x_with_bias, residual, prob, True)


def bias_dropout_add_fused_inference(x_with_bias, residual, prob):
    return _bda(x_with_bias, residual, prob, False)


def get_bias_dropout_add(training: bool, fused: bool) -> Callable:
    """Both variants run the same kernel here (there is no slower "unfused" implementation to select)."""
    return bias_dropout_add_fused_train if training else bias_dropout_add_fused_inference


# -- fused_softmax -------------------------------------------------------------------------------------------------------
class ScaledUpperTriangMaskedSoftmax(torch.autograd.Function):
    """softmax(scale * x) under a causal mask over ``[attn_batches, sq, sk]``."""

    @staticmethod
    def forward(ctx, inputs, scale):
        sq, sk = inputs.shape[-2:]
        mask = torch.ones(sq, sk, dtype=torch.bool, device=inputs.device).triu(1)
        probs = torch.softmax((inputs.float() * scale).masked_fill(mask, float("-inf")), dim=-1).to(inputs.dtype)
        ctx.save_for_backward(probs)
        ctx.scale = scale
        return probs

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        pf, gf = p.float(), g.float()
        return ((gf - (gf * pf).sum(-1, keepdim=True)) * pf * ctx.scale).to(g.dtype), None


class ScaledMaskedSoftmax(torch.autograd.Function):
    """softmax(scale * x) with a boolean mask (True = masked) broadcastable to ``[b, np, sq, sk]``."""

    @staticmethod
    def forward(ctx, inputs, mask, scale):
        x = inputs.float() * scale
        if mask is not None:
            x = x.masked_fill(mask, -10000.0)
        probs = torch.softmax(x, dim=-1).to(inputs.dtype)
        ctx.save_for_backward(probs)
        ctx.scale = scale
        return probs

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        pf, gf = p.float(), g.float()
        return ((gf - (gf * pf).sum(-1, keepdim=True)) * pf * ctx.scale).to(g.dtype), None, None


class ScaledSoftmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, scale):
        return ScaledMaskedSoftmax.forward(ctx, inputs, None, scale)

    @staticmethod
    def backward(ctx, g):
        return ScaledMaskedSoftmax.backward(ctx, g)[0], None


class FusedScaleMaskSoftmax(torch.nn.Module):
    """scale -> mask -> softmax of attention scores ``[b, np, sq, sk]`` (reference ``fusions/fused_softmax.py``).  Only the
    un-fused attention path needs it: the flash-attention kernel computes the softmax on chip."""

    def __init__(self, input_in_fp16: bool, input_in_bf16: bool, attn_mask_type: AttnMaskType, scaled_masked_softmax_fusion: bool,
                 mask_func: Optional[Callable], softmax_in_fp32: bool, scale: Optional[float]):
        super().__init__()
        if input_in_fp16 and input_in_bf16:
            raise ValueError("both fp16 and bf16 flags cannot be active at the same time")
        if scale is not None and not softmax_in_fp32:
            raise ValueError("softmax should be in fp32 when scaled")
        self.input_in_float16 = input_in_fp16 or input_in_bf16
        self.attn_mask_type, self.mask_func, self.softmax_in_fp32, self.scale = attn_mask_type, mask_func, softmax_in_fp32, scale
        self.scaled_masked_softmax_fusion = scaled_masked_softmax_fusion

    def forward(self, input: torch.Tensor, mask: Optional[torch.Tensor]):
        scale = self.scale if self.scale is not None else 1.0
        if self.attn_mask_type == AttnMaskType.causal and mask is None:
            b, np_, sq, sk = input.shape
            return ScaledUpperTriangMaskedSoftmax.apply(input.reshape(-1, sq, sk), scale).view(b, np_, sq, sk)
        if self.mask_func is not None and mask is not None:
            x = input.float() if (self.input_in_float16 and self.softmax_in_fp32) else input
            probs = torch.softmax(self.mask_func(x * scale, mask), dim=-1)
            return probs.to(input.dtype)
        return ScaledMaskedSoftmax.apply(input, mask, scale)


# -- fused_cross_entropy ---------------------------------------------------------------------------------------------------
def fused_vocab_parallel_cross_entropy(vocab_parallel_logits, target, tp_group=None):
    """Per-token loss over vocabulary shards with ONE packed all-reduce pair (``ops.cross_entropy``)."""
    from ..parallel.cross_entropy import vocab_parallel_cross_entropy

    return vocab_parallel_cross_entropy(vocab_parallel_logits, target)


# -- remaining function names of fused_bias_geglu / fused_bias_swiglu / fused_cross_entropy ---------------------------------------------
def geglu_back(g, y):
    def dgelu(a):
        t = torch.tanh(0.79788456 * a * (1 + 0.044715 * a * a))
        return 0.5 * a * ((1 - t * t) * (0.79788456 + 0.1070322243 * a * a)) + 0.5 * (1 + t)

    return _glu_back(g, y, lambda a: F.gelu(a, approximate="tanh"), dgelu).to(y.dtype)


def bias_geglu_back(g, y, bias):
    return geglu_back(g, y + bias)


class GeGLUFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return geglu(input)

    @staticmethod
    def backward(ctx, grad_output):
        (input,) = ctx.saved_tensors
        return geglu_back(grad_output, input)


class BiasGeGLUFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, bias):
        ctx.save_for_backward(input, bias)
        return bias_geglu(bias, input)

    @staticmethod
    def backward(ctx, grad_output):
        input, bias = ctx.saved_tensors
        tmp = bias_geglu_back(grad_output, input, bias)
        return tmp, tmp.reshape(-1, tmp.shape[-1]).sum(0)


def weighted_swiglu(y, weights):
    return (ops.swiglu(y, None).float() * weights.float()).to(y.dtype)


def weighted_swiglu_back(g, y, weights):
    """-> (d input, d weights) of ``weighted_swiglu``."""
    out = ops.swiglu(y.detach(), None).float()
    gin = swiglu_back(g.float() * weights.float(), y)
    return gin, (out * g.float()).sum(dim=-1, keepdim=True).to(weights.dtype)


class WeightedSwiGLUFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, weights, fp8_input_store=False):
        ctx.save_for_backward(input, weights)
        return weighted_swiglu(input, weights)

    @staticmethod
    def backward(ctx, grad_output):
        input, weights = ctx.saved_tensors
        gin, gw = weighted_swiglu_back(grad_output, input, weights)
        return gin, gw, None


def calculate_logits_max(vocab_parallel_logits):
    """fp32 logits and their row maximum over this vocabulary shard (to be MAX-all-reduced)."""
    logits = vocab_parallel_logits.float()
    return logits, torch.max(logits, dim=-1)[0]


def calculate_predicted_logits(vocab_parallel_logits, target, logits_max, vocab_start_index: int, vocab_end_index: int):
    """Shifted logits, this shard's contribution to the target logit and to the soft-max normaliser (both to be SUM-all-reduced)."""
    logits = vocab_parallel_logits - logits_max.unsqueeze(dim=-1)
    target_mask = (target < vocab_start_index) | (target >= vocab_end_index)
    masked_target = (target - vocab_start_index).masked_fill(target_mask, 0)
    flat = logits.reshape(-1, logits.shape[-1])
    picked = flat[torch.arange(flat.shape[0], device=flat.device), masked_target.reshape(-1)].view_as(target)
    predicted = picked.masked_fill(target_mask, 0.0)
    exp_logits = torch.exp(logits)
    return target_mask, masked_target.reshape(-1), predicted, exp_logits.sum(dim=-1), exp_logits


def calculate_cross_entropy_loss(exp_logits, predicted_logits, sum_exp_logits):
    """-> (softmax probabilities of this shard, per-token loss) from the reduced sums."""
    return exp_logits / sum_exp_logits.unsqueeze(dim=-1), torch.log(sum_exp_logits) - predicted_logits


def calculate_gradients(softmax, grad_output, target_mask, masked_target_1d):
    grad = softmax.clone().reshape(-1, softmax.shape[-1])
    rows = torch.arange(grad.shape[0], device=grad.device)
    grad[rows, masked_target_1d] -= 1.0 - target_mask.reshape(-1).float()
    return grad.view_as(softmax) * grad_output.unsqueeze(dim=-1)
