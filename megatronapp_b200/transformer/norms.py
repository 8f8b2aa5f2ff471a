"""LayerNorm / RMSNorm / L2Norm modules backed by the fused sm_100a norm kernel.

Parity: reference ``fusions/fused_layer_norm.py:30-169`` (``FusedLayerNorm``, zero-centred gamma,
``sequence_parallel`` attribute on the params) and ``transformer/torch_norm.py:9-96``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class FusedLayerNorm(nn.Module):
    def __init__(self, config, hidden_size: int, eps: float = 1e-5, persist_layer_norm: bool = True,
                 zero_centered_gamma: bool = False, normalization: str = "LayerNorm"):
        super().__init__()
        self.config = config
        self.zero_centered_gamma = config.layernorm_zero_centered_gamma if config is not None else zero_centered_gamma
        self.normalization = config.normalization if config is not None else normalization
        if self.normalization not in ("LayerNorm", "RMSNorm"):
            raise ValueError(f"unsupported normalization {self.normalization}")
        self.eps = eps
        self.hidden_size = hidden_size
        dtype = config.params_dtype if config is not None else torch.float32
        dev = "cpu" if (config is None or config.use_cpu_initialization or not torch.cuda.is_available()) \
            else torch.cuda.current_device()
        init = 0.0 if self.zero_centered_gamma else 1.0
        self.weight = nn.Parameter(torch.full((hidden_size,), init, dtype=dtype, device=dev))
        if self.normalization == "LayerNorm":
            self.bias = nn.Parameter(torch.zeros(hidden_size, dtype=dtype, device=dev))
        else:
            self.register_parameter("bias", None)
        sp = bool(config.sequence_parallel) if config is not None else False
        setattr(self.weight, "sequence_parallel", sp)
        if self.bias is not None:
            setattr(self.bias, "sequence_parallel", sp)

    def forward(self, x):
        return ops.layer_norm(x, self.weight, self.bias, self.eps, rms=self.normalization == "RMSNorm",
                              zero_centered=self.zero_centered_gamma)


WrappedTorchNorm = FusedLayerNorm
LNImpl = FusedLayerNorm


class L2Norm(nn.Module):
    def __init__(self, hidden_size: int = 0, eps: float = 1e-6, **_):
        super().__init__()
        self.eps = eps

    def forward(self, x):
        xf = x.float()
        return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)).type_as(x)


class IdentityOp(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, x, *a, **k):
        return x


class IdentityFuncOp(IdentityOp):
    def forward(self, *a, **k):
        return super().forward
