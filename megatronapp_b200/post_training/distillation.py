"""Knowledge-distillation losses for a teacher/student pair sharing the tensor-parallel layout.

Parity: reference ``megatron/post_training/`` + ``core/post_training/modelopt`` distillation pieces
(``LogitsKLLoss`` with temperature over vocab-parallel logits -- softmax normalisers all-reduced over
the TP group --, ``HiddenStateCosineLoss`` on intermediate activations, and the balancer that mixes
the KD terms with the original LM loss ``--export-kd-*`` style).  Quantisation (PTQ) itself relies on
the external ModelOpt package and is not reproduced.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from ..parallel import state as ps


class _AllReduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, op, group):
        y = x.clone()
        dist.all_reduce(y, op=op, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None, None


def _tp_reduce(x, op):
    g = ps.get_tensor_model_parallel_group()
    if ps.get_tensor_model_parallel_world_size() > 1:
        return _AllReduce.apply(x, op, g)
    return x


class LogitsKLLoss(torch.nn.Module):
    """KL(teacher || student) over vocab-parallel logits ``[s, b, V/tp]`` with temperature."""

    def __init__(self, temperature: float = 1.0, reverse: bool = False):
        super().__init__()
        self.t, self.reverse = temperature, reverse

    def forward(self, student_logits, teacher_logits):
        s = student_logits.float() / self.t
        te = teacher_logits.float().detach() / self.t

        def log_softmax_tp(x):
            m = _tp_reduce(x.max(dim=-1, keepdim=True).values.detach(), dist.ReduceOp.MAX)
            z = x - m
            lse = torch.log(_tp_reduce(z.exp().sum(dim=-1, keepdim=True), dist.ReduceOp.SUM))
            return z - lse

        ls, lt = log_softmax_tp(s), log_softmax_tp(te)
        if self.reverse:
            kl = (ls.exp() * (ls - lt)).sum(dim=-1)
        else:
            kl = (lt.exp() * (lt - ls)).sum(dim=-1)
        kl = _tp_reduce(kl, dist.ReduceOp.SUM)
        return kl * (self.t ** 2)             # [s, b]


class HiddenStateCosineLoss(torch.nn.Module):
    def forward(self, student_hidden, teacher_hidden):
        return 1.0 - F.cosine_similarity(student_hidden.float(), teacher_hidden.float().detach(), dim=-1)


class DistillationLossBalancer(torch.nn.Module):
    """total = (1 - kd_weight) * lm_loss + kd_weight * sum_i w_i * kd_i  (``skip_lm_loss`` drops the first term)."""

    def __init__(self, kd_loss_scale: float = 1.0, skip_original_loss: bool = False, weights: Optional[Dict[str, float]] = None):
        super().__init__()
        self.kd_loss_scale, self.skip, self.weights = kd_loss_scale, skip_original_loss, weights or {}

    def forward(self, lm_loss, kd_losses: Dict[str, torch.Tensor]):
        kd = sum(self.weights.get(k, 1.0) * v.mean() for k, v in kd_losses.items())
        if self.skip or lm_loss is None:
            return kd * self.kd_loss_scale
        lm = lm_loss.mean()
        # scale the KD term to the magnitude of the LM loss so neither dominates (ModelOpt balancer)
        scale = (lm.detach() / kd.detach().clamp_min(1e-8)) if self.kd_loss_scale == 0 else self.kd_loss_scale
        return lm + kd * scale


@torch.no_grad()
def _teacher_forward(teacher, *a, **k):
    teacher.eval()
    return teacher(*a, **k)


def distillation_forward(student, teacher, tokens, position_ids, attention_mask, labels, loss_mask,
                         logits_loss: Optional[LogitsKLLoss] = None, balancer: Optional[DistillationLossBalancer] = None):
    """One KD step: returns (total_loss, {"lm loss":..., "kd loss":...})."""
    logits_loss = logits_loss or LogitsKLLoss()
    balancer = balancer or DistillationLossBalancer()
    s_logits = student(tokens, position_ids, attention_mask)                     # [b, s, V/tp]
    t_logits = _teacher_forward(teacher, tokens, position_ids, attention_mask)
    kd = logits_loss(s_logits, t_logits)
    mask = loss_mask.float()
    kd = (kd * mask).sum() / mask.sum().clamp_min(1)
    lm = None
    if labels is not None:
        from ..parallel.cross_entropy import vocab_parallel_cross_entropy

        ce = vocab_parallel_cross_entropy(s_logits.transpose(0, 1).contiguous().float(), labels.transpose(0, 1).contiguous())
        lm = (ce.transpose(0, 1) * mask).sum() / mask.sum().clamp_min(1)
    total = balancer(lm, {"logits": kd})
    return total, {"lm loss": lm.detach() if lm is not None else None, "kd loss": kd.detach()}
