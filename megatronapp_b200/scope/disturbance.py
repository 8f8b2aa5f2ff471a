"""MegaScope perturbation injection (reference ``megatron/core/tensor_disturbance.py:27-74``):
three injection points (weight / calculation / system) x two noise functions
(``noise1``: x + N(0,1)*coef, ``noise2``: x * U(1-coef, 1+coef)); config keys
``{weight,calculation,system}_perturbation[_fn|_coef]``."""
from __future__ import annotations

from typing import Any, Dict

import torch


def noise1(x: torch.Tensor, coef: float) -> torch.Tensor:
    return x + torch.randn_like(x) * coef


def noise2(x: torch.Tensor, coef: float) -> torch.Tensor:
    return x * torch.empty_like(x).uniform_(1.0 - coef, 1.0 + coef)


_FNS = {"noise1": noise1, "noise2": noise2}


class Disturbance:
    def __init__(self):
        self.weight_perturbation = self.calculation_perturbation = self.system_perturbation = False
        self.fn = {"weight": "noise1", "calculation": "noise1", "system": "noise1"}
        self.coef = {"weight": 0.0, "calculation": 0.0, "system": 0.0}

    def set_by_configs(self, configs: Dict[str, Any]):
        for kind in ("weight", "calculation", "system"):
            on = configs.get(f"{kind}_perturbation", False)
            on = on if isinstance(on, bool) else str(on).lower() == "true"
            setattr(self, f"{kind}_perturbation", on)
            fn = str(configs.get(f"{kind}_perturbation_fn", self.fn[kind]))
            if fn not in _FNS:
                raise ValueError(f"unknown perturbation function {fn!r}; choose noise1 | noise2")
            self.fn[kind] = fn
            self.coef[kind] = float(configs.get(f"{kind}_perturbation_coef", self.coef[kind]))

    def apply(self, kind: str, x: torch.Tensor) -> torch.Tensor:
        if kind == "weight":
            with torch.no_grad():
                return _FNS[self.fn[kind]](x, self.coef[kind])
        return _FNS[self.fn[kind]](x, self.coef[kind])


def get_disturbance() -> "Disturbance":
    """The installed perturbation object (created on first use), reference ``tensor_disturbance.get_disturbance``."""
    from . import hooks

    if hooks._DISTURB is None:
        hooks.install(hooks._CAPTURE, Disturbance())
    return hooks._DISTURB


def noise1_factory(coef: float):
    return lambda x: noise1(x, coef)


def noise2_factory(val: float):
    return lambda x: noise2(x, val)
