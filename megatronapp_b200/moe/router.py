"""Top-k router with load-balancing (aux / seq-aux / sinkhorn / none), z-loss, jitter, expert bias.

Parity: reference ``transformer/moe/router.py:102-439`` (``TopKRouter``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..parallel import state as ps
from ..parallel.mappings import gather_from_sequence_parallel_region
from ..transformer.module import MegatronModule
from .moe_utils import (MoEAuxLossAutoScaler, save_to_aux_losses_tracker, sequence_load_balancing_loss_func,
                        switch_load_balancing_loss_func, topk_softmax_with_capacity, z_loss_func)


def sinkhorn(cost, tol=1e-4):
    cost = torch.exp(cost)
    d0 = torch.ones(cost.size(0), device=cost.device, dtype=cost.dtype)
    d1 = torch.ones(cost.size(1), device=cost.device, dtype=cost.dtype)
    eps, err, d1_old = 1e-8, 1e9, d1
    while err > tol:
        d0 = (1 / d0.size(0)) / (torch.sum(d1 * cost, 1) + eps)
        d1 = (1 / d1.size(0)) / (torch.sum(d0.unsqueeze(1) * cost, 0) + eps)
        err = torch.mean(torch.abs(d1_old - d1))
        d1_old = d1
    return d1 * cost * d0.unsqueeze(1)


class TopKRouter(MegatronModule):
    def __init__(self, config):
        super().__init__(config)
        self.num_experts = config.num_moe_experts
        self.topk = config.moe_router_topk
        self.routing_type = config.moe_router_load_balancing_type
        self.score_function = config.moe_router_score_function
        self.input_jitter = None
        self.layer_number = None
        dev = "cpu" if (config.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        self.weight = nn.Parameter(torch.empty(self.num_experts, config.hidden_size, dtype=torch.float32, device=dev))
        if config.perform_initialization:
            config.init_method(self.weight)
        self.weight.data = self.weight.data.to(dtype=config.params_dtype)
        setattr(self.weight, "sequence_parallel", config.sequence_parallel)
        if config.moe_router_enable_expert_bias:
            self.register_buffer("local_tokens_per_expert", torch.zeros(self.num_experts, dtype=torch.float32, device=dev), persistent=False)
            self.register_buffer("expert_bias", torch.zeros(self.num_experts, dtype=torch.float32, device=dev))
        else:
            self.local_tokens_per_expert = None
            self.expert_bias = None

    def set_layer_number(self, n):
        self.layer_number = n

    def gating(self, x):
        dt = torch.float32 if self.config.moe_router_dtype == "fp32" else (
            torch.float64 if self.config.moe_router_dtype == "fp64" else x.dtype)
        return torch.nn.functional.linear(x.to(dt), self.weight.to(dt))

    def apply_input_jitter(self, x):
        eps = self.config.moe_input_jitter_eps
        if eps is None or not self.training:
            return x
        if self.input_jitter is None:
            self.input_jitter = torch.distributions.uniform.Uniform(
                torch.tensor(1.0 - eps, device=x.device), torch.tensor(1.0 + eps, device=x.device)).rsample
        return x * self.input_jitter(x.shape).to(x.dtype)

    def apply_z_loss(self, logits):
        c = self.config.moe_z_loss_coeff
        if c is not None and self.training:
            z = z_loss_func(logits, c / ps.get_tensor_and_context_parallel_world_size())
            logits = MoEAuxLossAutoScaler.apply(logits, z)
            save_to_aux_losses_tracker("z_loss", z / c, self.layer_number, self.config.num_layers)
        return logits

    def routing(self, logits):
        seq_length, bsz = logits.shape[:2]
        logits = logits.view(-1, self.num_experts)
        logits = self.apply_z_loss(logits)
        cfg = self.config
        if self.routing_type == "sinkhorn":
            with torch.no_grad():
                norm = sinkhorn(logits.to(dtype=torch.float32))
                _, idx = torch.topk(norm, k=self.topk, dim=1)
            act = torch.sigmoid(logits) if self.topk == 1 else torch.softmax(logits, dim=-1, dtype=torch.float32).type_as(logits)
            rmap = torch.zeros_like(logits).int().scatter(1, idx, 1).bool()
            return act * rmap, rmap
        fused_ok = (logits.is_cuda and self.score_function == "softmax" and cfg.moe_expert_capacity_factor is None
                    and not cfg.moe_router_group_topk and self.expert_bias is None and self.num_experts <= 256
                    and self.topk <= 8 and logits.dtype in (torch.float32, torch.bfloat16, torch.float16))
        if fused_ok:
            # one warp per token: softmax + top-k (+ renormalisation / scaling), dense probs + routing map out
            from ..ops.moe import router_topk

            probs, rmap, tpe = router_topk(logits, self.topk, bool(cfg.moe_router_pre_softmax),
                                           cfg.moe_router_topk_scaling_factor)
        else:
            probs, rmap, tpe = self._routing_torch(logits, cfg)
        if self.training and cfg.moe_aux_loss_coeff > 0 and self.routing_type in ("aux_loss", "seq_aux_loss"):
            scores = torch.softmax(logits, dim=-1, dtype=torch.float32) if self.score_function == "softmax" else (
                lambda s: s / (s.sum(-1, keepdim=True) + 1e-20))(torch.sigmoid(logits.float()))
            coeff = cfg.moe_aux_loss_coeff / ps.get_tensor_and_context_parallel_world_size()
            if self.routing_type == "aux_loss":
                aux = switch_load_balancing_loss_func(scores, tpe.float(), self.topk, coeff)
            else:
                aux = sequence_load_balancing_loss_func(scores, rmap, bsz, seq_length, self.topk, coeff)
            save_to_aux_losses_tracker("load_balancing_loss", aux / cfg.moe_aux_loss_coeff, self.layer_number, cfg.num_layers)
            probs = MoEAuxLossAutoScaler.apply(probs, aux)
        if self.local_tokens_per_expert is not None and self.training:
            with torch.no_grad():
                self.local_tokens_per_expert += rmap.sum(dim=0)
        return probs, rmap

    def _routing_torch(self, logits, cfg):
        return topk_softmax_with_capacity(
            logits, self.topk, cfg.moe_expert_capacity_factor, cfg.moe_pad_expert_input_to_capacity,
            cfg.moe_token_drop_policy, cfg.moe_router_pre_softmax, cfg.moe_router_num_groups,
            cfg.moe_router_group_topk, cfg.moe_router_topk_scaling_factor, cfg.deterministic_mode,
            self.score_function, self.expert_bias)

    def forward(self, x):
        x = self.apply_input_jitter(x)
        logits = self.gating(x)
        return self.routing(logits)
