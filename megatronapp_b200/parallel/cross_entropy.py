"""Vocab-parallel cross entropy.

``vocab_parallel_cross_entropy(logits[s,b,V/tp], target[s,b]) -> loss[s,b]``; API parity with
reference ``tensor_parallel/cross_entropy.py:123-233``.  On CUDA the statistics / gradient run
in the fused sm_100a kernels (``ops.cross_entropy``: one packed all-reduce after the MAX,
softmax recomputed in backward, gradient written in place of the logits).  The PyTorch path
below is the fp32 reference used on CPU and by the numerics tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import state as ps


class _VocabParallelCrossEntropyRef(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, vocab_start, vocab_total, group, label_smoothing):
        lf = logits.float()
        tp = group is not None and dist.is_initialized() and not isinstance(group, ps._SingleGroup) \
            and dist.get_world_size(group) > 1
        mx = lf.max(dim=-1).values
        if tp:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        lf = lf - mx.unsqueeze(-1)
        vl = lf.shape[-1]
        local = target - vocab_start
        mask = (local < 0) | (local >= vl)
        local = local.masked_fill(mask, 0)
        tgt = lf.gather(-1, local.unsqueeze(-1)).squeeze(-1).masked_fill(mask, 0.0)
        ex = lf.exp()
        se = ex.sum(-1)
        if tp:
            packed = torch.stack([se, tgt])
            dist.all_reduce(packed, group=group)
            se, tgt = packed[0], packed[1]
        loss = se.log() - tgt
        softmax = ex / se.unsqueeze(-1)
        if label_smoothing > 0:
            k = label_smoothing * vocab_total / (vocab_total - 1)
            logp_sum = (lf - se.log().unsqueeze(-1)).sum(-1)
            if tp:
                dist.all_reduce(logp_sum, group=group)
            loss = (1.0 - k) * loss - k * logp_sum / vocab_total
        ctx.save_for_backward(softmax, mask, local)
        ctx.ls, ctx.vocab_total, ctx.dtype = label_smoothing, vocab_total, logits.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        softmax, mask, local = ctx.saved_tensors
        grad = softmax.clone()
        ls, V = ctx.ls, ctx.vocab_total
        upd = (1.0 - mask.float())
        if ls > 0:
            k = ls * V / (V - 1)
            grad.scatter_add_(-1, local.unsqueeze(-1), (-(1.0 - k) * upd).unsqueeze(-1))
            grad -= k / V
        else:
            grad.scatter_add_(-1, local.unsqueeze(-1), (-upd).unsqueeze(-1))
        grad = grad * g.unsqueeze(-1)
        return grad.to(ctx.dtype), None, None, None, None, None


def vocab_parallel_cross_entropy_ref(logits, target, vocab_start=0, vocab_total=None, group=None,
                                     label_smoothing=0.0):
    vocab_total = vocab_total or logits.shape[-1]
    return _VocabParallelCrossEntropyRef.apply(logits, target, vocab_start, vocab_total, group, label_smoothing)


def vocab_parallel_cross_entropy(vocab_parallel_logits, target, label_smoothing: float = 0.0):
    from .. import ops

    tp = ps.get_tensor_model_parallel_world_size()
    rank = ps.get_tensor_model_parallel_rank()
    vl = vocab_parallel_logits.shape[-1]
    group = ps.get_tensor_model_parallel_group(check_initialized=False) if tp > 1 else None
    return ops.cross_entropy(vocab_parallel_logits, target, vocab_start=rank * vl, vocab_total=vl * tp,
                             group=group, label_smoothing=label_smoothing)
