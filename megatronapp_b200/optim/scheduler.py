"""Learning-rate / weight-decay schedules (linear, cosine, inverse-square-root, WSD) with warmup.

Parity: reference ``core/optimizer_param_scheduler.py:1-297`` (checkpointable state, override /
use-checkpoint semantics, ``wsd_decay_steps`` + ``lr_wsd_decay_style``).
"""
from __future__ import annotations

import math
from typing import Optional


class OptimizerParamScheduler:
    def __init__(self, optimizer, init_lr: float, max_lr: float, min_lr: float, lr_warmup_steps: int,
                 lr_decay_steps: int, lr_decay_style: str, start_wd: float, end_wd: float, wd_incr_steps: int,
                 wd_incr_style: str, use_checkpoint_opt_param_scheduler: bool = True,
                 override_opt_param_scheduler: bool = False, wsd_decay_steps: Optional[int] = None,
                 lr_wsd_decay_style: Optional[str] = None):
        self.optimizer = optimizer
        self.init_lr, self.max_lr, self.min_lr = init_lr, float(max_lr), min_lr
        if not (0.0 <= self.min_lr <= self.max_lr and self.init_lr <= self.max_lr):
            raise ValueError("need 0 <= min_lr <= max_lr and init_lr <= max_lr")
        self.lr_warmup_steps, self.lr_decay_steps = lr_warmup_steps, lr_decay_steps
        if lr_decay_steps <= 0 or lr_warmup_steps >= lr_decay_steps:
            raise ValueError("need 0 < lr_decay_steps and lr_warmup_steps < lr_decay_steps")
        self.lr_decay_style = lr_decay_style
        self.wsd_decay_steps, self.lr_wsd_decay_style = wsd_decay_steps, lr_wsd_decay_style
        if lr_decay_style == "WSD" and wsd_decay_steps is None:
            raise ValueError("WSD needs wsd_decay_steps")
        self.start_wd, self.end_wd = start_wd, end_wd
        if not (0.0 <= start_wd <= end_wd):
            raise ValueError("need 0 <= start_wd <= end_wd")
        self.wd_incr_steps, self.wd_incr_style = wd_incr_steps, wd_incr_style
        self.override_opt_param_scheduler = override_opt_param_scheduler
        self.use_checkpoint_opt_param_scheduler = use_checkpoint_opt_param_scheduler
        if override_opt_param_scheduler and use_checkpoint_opt_param_scheduler:
            raise ValueError("both override and use-checkpoint are set")
        self.num_steps = 0
        self.step(0)

    def get_wd(self) -> float:
        if self.num_steps > self.wd_incr_steps:
            return self.end_wd
        if self.wd_incr_style == "constant":
            if self.start_wd != self.end_wd:
                raise ValueError("constant wd needs start_wd == end_wd")
            return self.end_wd
        ratio = float(self.num_steps) / float(self.wd_incr_steps)
        delta = self.end_wd - self.start_wd
        if self.wd_incr_style == "linear":
            coeff = ratio
        elif self.wd_incr_style == "cosine":
            coeff = 0.5 * (math.cos(math.pi * (1 - ratio)) + 1.0)
        else:
            raise ValueError(f"{self.wd_incr_style} weight decay increment style is not supported")
        return self.start_wd + coeff * delta

    def get_lr(self, param_group: dict) -> float:
        max_lr = param_group.get("max_lr", self.max_lr)
        min_lr = param_group.get("min_lr", self.min_lr)
        if self.lr_warmup_steps > 0 and self.num_steps <= self.lr_warmup_steps:
            return self.init_lr + (max_lr - self.init_lr) * float(self.num_steps) / float(self.lr_warmup_steps)
        if self.lr_decay_style == "constant":
            return max_lr
        if self.num_steps > self.lr_decay_steps:
            return min_lr
        if self.lr_decay_style == "inverse-square-root":
            warm = max(self.lr_warmup_steps, 1)
            return max(min_lr, max_lr * warm ** 0.5 / max(self.num_steps, 1) ** 0.5)
        num = self.num_steps - self.lr_warmup_steps
        den = self.lr_decay_steps - self.lr_warmup_steps
        ratio = float(num) / float(den)
        delta = max_lr - min_lr
        if self.lr_decay_style == "linear":
            coeff = 1.0 - ratio
        elif self.lr_decay_style == "cosine":
            coeff = 0.5 * (math.cos(math.pi * ratio) + 1.0)
        elif self.lr_decay_style == "WSD":
            anneal_start = self.lr_decay_steps - self.wsd_decay_steps
            if self.num_steps <= anneal_start:
                coeff = 1.0
            else:
                r = float(self.num_steps - anneal_start) / float(self.wsd_decay_steps)
                style = self.lr_wsd_decay_style or "exponential"
                if style == "linear":
                    coeff = 1.0 - r
                elif style == "cosine":
                    coeff = 0.5 * (math.cos(math.pi * r) + 1.0)
                elif style == "exponential":
                    coeff = (2.0 * math.pow(0.5, r)) - 1.0
                elif style == "minus_sqrt":
                    coeff = 1.0 - math.sqrt(r)
                else:
                    raise ValueError(style)
        else:
            raise ValueError(f"{self.lr_decay_style} decay style is not supported")
        return min_lr + coeff * delta

    def step(self, increment: int) -> None:
        self.num_steps += increment
        wd = self.get_wd()
        for g in self.optimizer.param_groups:
            g["lr"] = self.get_lr(g) * g.get("lr_mult", 1.0)
            g["weight_decay"] = wd * g.get("wd_mult", 1.0)

    def state_dict(self):
        return {k: getattr(self, k) for k in (
            "max_lr", "lr_warmup_steps", "num_steps", "lr_decay_style", "lr_decay_steps", "min_lr",
            "start_wd", "end_wd", "wd_incr_style", "wd_incr_steps")}

    def _check_and_set(self, cls_value, sd_value, name):
        if self.override_opt_param_scheduler:
            return cls_value
        if not self.use_checkpoint_opt_param_scheduler and cls_value != sd_value:
            raise ValueError(f"scheduler value for {name} ({cls_value}) differs from checkpoint ({sd_value})")
        return sd_value

    def load_state_dict(self, sd):
        self.max_lr = self._check_and_set(self.max_lr, sd.get("max_lr", sd.get("start_lr")), "max_lr")
        self.min_lr = self._check_and_set(self.min_lr, sd["min_lr"], "min_lr")
        self.lr_warmup_steps = self._check_and_set(self.lr_warmup_steps, sd.get("lr_warmup_steps", sd.get("warmup_iter")), "warmup")
        self.lr_decay_steps = self._check_and_set(self.lr_decay_steps, sd.get("lr_decay_steps", sd.get("end_iter")), "decay steps")
        self.lr_decay_style = self._check_and_set(self.lr_decay_style, sd.get("lr_decay_style", sd.get("decay_style")), "decay style")
        n = sd.get("num_steps", sd.get("num_iters", 0))
        self.num_steps = 0
        self.step(increment=n)
        if "start_wd" in sd:
            self.start_wd = self._check_and_set(self.start_wd, sd["start_wd"], "start_wd")
            self.end_wd = self._check_and_set(self.end_wd, sd["end_wd"], "end_wd")
            self.wd_incr_steps = self._check_and_set(self.wd_incr_steps, sd["wd_incr_steps"], "wd steps")
            self.wd_incr_style = self._check_and_set(self.wd_incr_style, sd["wd_incr_style"], "wd style")
