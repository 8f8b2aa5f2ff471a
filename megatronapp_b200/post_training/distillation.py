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
    """all-reduce whose backward is either the identity (``reduce_grad=False``: the result feeds a loss every rank holds a full
    copy of) or another all-reduce (``reduce_grad=True``: the result feeds rank-local terms, so the cotangents of all ranks add)."""

    @staticmethod
    def forward(ctx, x, op, group, reduce_grad):
        ctx.group, ctx.reduce_grad = group, reduce_grad
        y = x.clone()
        dist.all_reduce(y, op=op, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        if ctx.reduce_grad:
            g = g.clone()
            dist.all_reduce(g, group=ctx.group)
        return g, None, None, None


def _tp_reduce(x, op, reduce_grad=False):
    g = ps.get_tensor_model_parallel_group()
    if ps.get_tensor_model_parallel_world_size() > 1:
        return _AllReduce.apply(x, op, g, reduce_grad)
    return x


def _first(x):
    return x[0] if isinstance(x, tuple) else x        # the parallel linears return (output, bias)


class LogitsKLLoss(torch.nn.Module):
    """KL(teacher || student) over vocab-parallel logits ``[s, b, V/tp]`` with temperature -> ``[s, b]``.

    Every rank holds a vocabulary slice: the soft-max normalisers are all-reduced (their cotangents too -- the normaliser enters
    the local KL terms of every rank), the per-token KL terms are summed over the group at the end (identity backward)."""

    def __init__(self, temperature: float = 1.0, reverse: bool = False):
        super().__init__()
        self.t, self.reverse = temperature, reverse

    def forward(self, student_logits, teacher_logits):
        s = _first(student_logits).float() / self.t
        te = _first(teacher_logits).float().detach() / self.t

        def log_softmax_tp(x):
            m = _tp_reduce(x.max(dim=-1, keepdim=True).values.detach(), dist.ReduceOp.MAX)
            z = x - m
            lse = torch.log(_tp_reduce(z.exp().sum(dim=-1, keepdim=True), dist.ReduceOp.SUM, reduce_grad=True))
            return z - lse

        ls, lt = log_softmax_tp(s), log_softmax_tp(te)
        if self.reverse:
            kl = (ls.exp() * (ls - lt)).sum(dim=-1)
        else:
            kl = (lt.exp() * (lt - ls)).sum(dim=-1)
        kl = _tp_reduce(kl, dist.ReduceOp.SUM)
        return kl * (self.t ** 2)             # [s, b]


class ProjectionLayer(torch.nn.Module):
    """student hidden size -> teacher hidden size (identity when they agree); reference ``algos/distillation.py``."""

    def __init__(self, hidden_size_student: Optional[int], hidden_size_teacher: Optional[int], sequence_parallel: bool = False):
        super().__init__()
        if hidden_size_student is None or hidden_size_teacher is None or hidden_size_student == hidden_size_teacher:
            self.fit = torch.nn.Identity()
        else:
            self.fit = torch.nn.Linear(hidden_size_student, hidden_size_teacher)
            torch.nn.init.normal_(self.fit.weight, mean=0.0, std=0.01)
            torch.nn.init.zeros_(self.fit.bias)
            for p in self.fit.parameters():      # replicated over TP, fed sequence shards: grads are summed over the group
                p.sequence_parallel = sequence_parallel

    def forward(self, x):
        return self.fit(x.to(self.fit.weight.dtype) if isinstance(self.fit, torch.nn.Linear) else x)


class HiddenStateCosineLoss(torch.nn.Module):
    """1 - cos(student, teacher) per token over the full hidden dimension, ``[s, b, h] -> [s, b]``."""

    def __init__(self, hidden_size_student: Optional[int] = None, hidden_size_teacher: Optional[int] = None,
                 sequence_parallel: bool = False):
        super().__init__()
        self.projection = ProjectionLayer(hidden_size_student, hidden_size_teacher, sequence_parallel)

    def forward(self, student_hidden, teacher_hidden):
        s = self.projection(_first(student_hidden))
        return 1.0 - F.cosine_similarity(s.float(), _first(teacher_hidden).float().detach(), dim=-1)


class MSELoss(HiddenStateCosineLoss):
    def forward(self, student_hidden, teacher_hidden):
        s = self.projection(_first(student_hidden))
        return F.mse_loss(s.float(), _first(teacher_hidden).float().detach(), reduction="none").sum(dim=-1)


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


class LogitsAndIntermediatesLossBalancer(torch.nn.Module):
    """Reference ``LogitsAndIntermediatesLossBalancer``: the intermediate losses are rescaled to the magnitude of the logit loss
    (and the KD scale halved when there are any); with the LM loss kept, the KD sum is rescaled to the LM loss's magnitude."""

    def __init__(self, kd_loss_scale: float = 1.0, skip_original_loss: bool = False):
        super().__init__()
        self.kd_loss_scale, self.skip = kd_loss_scale, skip_original_loss

    def forward(self, student_loss, logits_loss, intermediate_losses):
        inter = sum(intermediate_losses) if intermediate_losses else None
        scale = self.kd_loss_scale
        if inter is not None and float(inter.detach()) > 0:
            inter = inter * (logits_loss.detach() / inter.detach())
            scale = scale / 2.0
        kd = logits_loss if inter is None else logits_loss + inter
        if self.skip or student_loss is None:
            return kd
        kd = kd * scale
        return student_loss + kd * (student_loss.detach() / kd.detach().clamp_min(1e-12))


def load_distillation_config(config_path: Optional[str], student_cfg, teacher_cfg, sequence_parallel: bool = False) -> dict:
    """``--export-kd-cfg`` yaml (reference ``load_distillation_config``): ``logit_layers: [student, teacher]``,
    ``intermediate_layer_pairs: [[student, teacher], ...]``, ``skip_lm_loss``, ``kd_loss_scale``; ``None`` -> logits only."""
    if config_path:
        import yaml

        with open(config_path) as f:
            cfg = yaml.safe_load(f)
    else:
        cfg = {}
    cfg = {"logit_layers": ["output_layer", "output_layer"], "intermediate_layer_pairs": [], "skip_lm_loss": True,
           "kd_loss_scale": 1.0, **cfg}
    hs, ht = student_cfg.hidden_size, teacher_cfg.hidden_size
    cfg["criterion"] = {tuple(cfg["logit_layers"]): LogitsKLLoss(temperature=float(cfg.get("temperature", 1.0)),
                                                                 reverse=bool(cfg.get("reverse_kl", False)))}
    for pair in cfg["intermediate_layer_pairs"]:
        loss_cls = MSELoss if cfg.get("intermediate_loss", "cosine") == "mse" else HiddenStateCosineLoss
        cfg["criterion"][tuple(pair)] = loss_cls(hs, ht, sequence_parallel)
    cfg["loss_balancer"] = LogitsAndIntermediatesLossBalancer(cfg["kd_loss_scale"], cfg["skip_lm_loss"])
    return cfg


class DistillationState:
    """Teacher + hooks attached to a student ``GPTModel`` (what ModelOpt's ``DistillationModel`` wrapper does in the reference,
    ``post_training/model_provider.py:222-247``).  The student object stays the model the trainer sees -- same parameters, same
    checkpoint keys -- the frozen teacher hangs off it un-registered, so it is neither optimised, reduced nor saved."""

    def __init__(self, student: torch.nn.Module, teacher: torch.nn.Module, cfg: dict, sequence_parallel: bool = False):
        self.teacher, self.cfg, self.sequence_parallel = teacher, cfg, sequence_parallel
        self.pairs = list(cfg["criterion"].items())
        self.logit_pair = tuple(cfg["logit_layers"])
        self._s_out: Dict[str, object] = {}
        self._t_out: Dict[str, object] = {}
        self._handles = []
        smods, tmods = dict(student.named_modules()), dict(teacher.named_modules())
        for (sn, tn), _ in self.pairs:
            if sn not in smods or tn not in tmods:
                raise KeyError(f"distillation layer pair ({sn}, {tn}): no such submodule in the "
                               f"{'student' if sn not in smods else 'teacher'}")
            self._handles.append(smods[sn].register_forward_hook(self._capture(self._s_out, sn)))
            self._handles.append(tmods[tn].register_forward_hook(self._capture(self._t_out, tn)))
        for p in teacher.parameters():
            p.requires_grad_(False)
        teacher.eval()
        # trainable projections live on the student so the optimizer / DDP / device moves see them
        projections = [loss.projection for _, loss in self.pairs
                       if isinstance(getattr(getattr(loss, "projection", None), "fit", None), torch.nn.Linear)]
        if projections:
            student.kd_projections = torch.nn.ModuleList(projections)

    @staticmethod
    def _capture(store, name):
        def hook(_module, _inputs, output):
            store[name] = output
        return hook

    @torch.no_grad()
    def teacher_forward(self, *args, **kwargs):
        self.teacher.eval()
        return self.teacher(*args, **kwargs)

    def compute_kd_loss(self, student_loss=None, loss_reduction_fn=None):
        """-> (total, logits_kd) scalars; ``loss_reduction_fn`` maps a per-token ``[b, s]`` loss to a scalar (masked mean)."""
        from ..parallel.mappings import gather_from_sequence_parallel_region

        reduce = loss_reduction_fn or (lambda x: x.mean())
        logits_loss, inter = None, []
        for (sn, tn), loss in self.pairs:
            per_token = loss(self._s_out[sn], self._t_out[tn])            # [s, b] (sequence shard for intermediate layers)
            if (sn, tn) != self.logit_pair and self.sequence_parallel and ps.get_tensor_model_parallel_world_size() > 1:
                per_token = gather_from_sequence_parallel_region(per_token, tensor_parallel_output_grad=False)
            value = reduce(per_token.transpose(0, 1).contiguous())
            if (sn, tn) == self.logit_pair:
                logits_loss = value
            else:
                inter.append(value)
        self._s_out.clear()
        self._t_out.clear()
        return self.cfg["loss_balancer"](student_loss, logits_loss, inter), logits_loss.detach()

    def detach(self, student):
        """``--export-kd-finalize``: drop the hooks and the projections, leaving the plain student."""
        for h in self._handles:
            h.remove()
        if hasattr(student, "kd_projections"):
            del student.kd_projections
        if "_kd" in student.__dict__:
            del student.__dict__["_kd"]


def attach_teacher(student, teacher, cfg: dict, sequence_parallel: bool = False) -> DistillationState:
    state = DistillationState(student, teacher, cfg, sequence_parallel)
    student.__dict__["_kd"] = state
    return state


_NEMO_KEYS = {"encoder_seq_length": "seq_length"}


def teacher_args_from_checkpoint(args, load_dir: str):
    """Arguments of the teacher: the student's, overridden by ``<load_dir>/model_config.yaml`` (NeMo-style keys as in the reference,
    ``post_training/model_provider.py:_load_teacher_model_config``) or, without that file, by the architecture recorded in the
    checkpoint itself."""
    import copy
    import os

    t = copy.copy(args)
    path = os.path.join(load_dir, "model_config.yaml")
    if not os.path.exists(path):
        from ..training.checkpointing import load_args_from_checkpoint

        t.export_kd_teacher_load = load_dir
        before = (t.hidden_size, t.num_attention_heads, t.kv_channels)
        t = load_args_from_checkpoint(t, load_arg="export_kd_teacher_load")
        if (t.hidden_size, t.num_attention_heads) != before[:2] and getattr(t, "kv_channels", None) == before[2]:
            t.kv_channels = t.hidden_size // t.num_attention_heads
        return t
    import yaml

    with open(path) as f:
        cfg = yaml.safe_load(f)
    for k, v in list(cfg.items()):
        if k in _NEMO_KEYS:
            cfg[_NEMO_KEYS[k]] = v
    if "bias" in cfg:
        cfg["add_bias_linear"] = bool(cfg["bias"])
    if cfg.get("activation") == "swiglu":
        cfg["swiglu"] = True
    if "position_embedding_type" in cfg and cfg["position_embedding_type"] is None:
        cfg["position_embedding_type"], cfg["add_position_embedding"] = "none", False
    if "share_embeddings_and_output_weights" in cfg:
        cfg["untie_embeddings_and_output_weights"] = not cfg["share_embeddings_and_output_weights"]
    if cfg.get("normalization") == "layernorm1p":
        cfg["normalization"], cfg["apply_layernorm_1p"] = "LayerNorm", True
    elif str(cfg.get("normalization", "")).lower() == "rmsnorm":
        cfg["normalization"] = "RMSNorm"
    elif str(cfg.get("normalization", "")).lower() == "layernorm":
        cfg["normalization"] = "LayerNorm"
    recompute_kv = "kv_channels" not in cfg and ("hidden_size" in cfg or "num_attention_heads" in cfg)
    for k, v in cfg.items():
        if not isinstance(v, dict):
            setattr(t, k, v)
    if recompute_kv:
        t.kv_channels = t.hidden_size // t.num_attention_heads
    if "ffn_hidden_size" not in cfg and "hidden_size" in cfg:
        t.ffn_hidden_size = (int((4 * t.hidden_size * 2 / 3) / 64) * 64) if getattr(t, "swiglu", False) else 4 * t.hidden_size
    return t
