"""Feed-forward block: column-parallel fc1 -> fused bias+activation -> row-parallel fc2.

Parity: reference ``transformer/mlp.py:32-371`` (``MLP``, ``MLPSubmodules``; gelu / geglu /
swiglu / squared-relu; per-token-scale variant for MoE) with the fused activation kernels of
``fusions/fused_bias_{gelu,swiglu}.py`` replaced by sm_100a elementwise kernels.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Union

import torch
import torch.nn.functional as F

from .. import ops
from ..scan.tracer import scope as _scope
from ..scope import hooks as scope_hooks
from .module import MegatronModule
from .spec_utils import ModuleSpec, build_module


@dataclass
class MLPSubmodules:
    linear_fc1: Union[ModuleSpec, type] = None
    linear_fc2: Union[ModuleSpec, type] = None


def squared_relu(x):
    return torch.pow(F.relu(x), 2)


def erf_gelu(x):
    """Exact GELU (``--onnx-safe``; reference legacy/model/utils.py:erf_gelu).  The default activation is the tanh form
    (``--openai-gelu`` and the reference's fused bias-GELU), computed in the GEMM epilogue."""
    return F.gelu(x, approximate="none")


class MLP(MegatronModule):
    def __init__(self, config, submodules: MLPSubmodules, is_expert: bool = False,
                 input_size: Optional[int] = None):
        super().__init__(config)
        self.input_size = input_size if input_size is not None else config.hidden_size
        ffn = config.moe_ffn_hidden_size if is_expert else config.ffn_hidden_size
        fc1_out = ffn * 2 if config.gated_linear_unit else ffn
        self.linear_fc1 = build_module(
            submodules.linear_fc1, self.input_size, fc1_out, config=config, init_method=config.init_method,
            gather_output=False, bias=config.add_bias_linear, skip_bias_add=True, is_expert=is_expert,
            tp_comm_buffer_name="fc1")
        self.activation_func = config.activation_func
        self.linear_fc2 = build_module(
            submodules.linear_fc2, ffn, config.hidden_size, config=config,
            init_method=config.output_layer_init_method, bias=config.add_bias_linear,
            input_is_parallel=True, skip_bias_add=True, is_expert=is_expert, tp_comm_buffer_name="fc2")
        self.layer_number = None

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        """Gated MLPs keep ``[gate_r; up_r]`` in each TP rank's fc1 shard, so the rank-major concatenation is NOT the
        global tensor: gate and up are exposed as two shards of the same key, gate rows first (global layout
        ``[gate_0..gate_{tp-1}; up_0..up_{tp-1}]``), as the reference's ``apply_swiglu_sharded_factory`` (mlp.py:300-371)."""
        from dataclasses import replace

        from .module import sharded_state_dict_of

        sd = {}
        for name, child in self.named_children():
            sub = sharded_state_dict_of(child, f"{prefix}{name}.", sharded_offsets, metadata)
            if name == "linear_fc1" and self.config.gated_linear_unit:
                split = {}
                for k, st in sub.items():
                    if not (hasattr(st, "local_shape") and st.data is not None and st.data.dim() >= 1
                            and k.endswith(("weight", "bias"))):
                        split[k] = st
                        continue
                    pre = st.prepend_axis_num
                    half = st.local_shape[0] // 2
                    frag = list(st.axis_fragmentations or (1,) * len(st.global_shape))
                    tp_frag, tp_idx = frag[pre], st.global_offset[pre] // st.local_shape[0] if st.local_shape[0] else 0
                    g_rows = half * tp_frag                      # rows of the global gate block
                    frag[pre] = 2 * tp_frag
                    for j, (nm, base) in enumerate((("gate", 0), ("up", g_rows))):
                        goff = list(st.global_offset)
                        goff[pre] = base + tp_idx * half
                        split[f"{k}#{nm}"] = replace(st, data=st.data.narrow(0, j * half, half),
                                                     local_shape=(half,) + tuple(st.local_shape[1:]),
                                                     global_offset=tuple(goff), axis_fragmentations=tuple(frag))
                sub = split
            sd.update(sub)
        return sd

    def _activation(self, x, bias):
        cfg = self.config
        if cfg.gated_linear_unit:
            if self.activation_func in (F.silu,):
                # Megatron packs [gate | up] along the last dim of each TP shard
                return ops.swiglu(x, bias)
            if bias is not None:
                x = x + bias
            g, u = torch.chunk(x, 2, dim=-1)
            return self.activation_func(g) * u
        if self.activation_func in (F.gelu,) or getattr(self.activation_func, "__name__", "") == "_gelu":
            return ops.bias_gelu(x, bias)
        if bias is not None:
            x = x + bias
        return self.activation_func(x)

    def forward(self, hidden_states, per_token_scale=None):
        with _scope("MLP.forward"):
            is_gelu = (not self.config.gated_linear_unit and
                       (self.activation_func in (F.gelu,) or getattr(self.activation_func, "__name__", "") == "_gelu"))
            if (is_gelu and hasattr(self.linear_fc1, "can_fuse_gelu") and self.linear_fc1.can_fuse_gelu(hidden_states)
                    and not scope_hooks.wants(4, self.layer_number)):
                inter, _ = self.linear_fc1(hidden_states, fused_gelu=True)   # bias + GELU in the GEMM epilogue
            else:
                inter, bias = self.linear_fc1(hidden_states)
                scope_hooks.tik_tensor(4, self.layer_number, inter)
                inter = self._activation(inter, bias)
            if per_token_scale is not None:
                inter = inter * per_token_scale.unsqueeze(-1).to(inter.dtype)
            out, out_bias = self.linear_fc2(inter)
            scope_hooks.tik_tensor(5, self.layer_number, out)
        return out, out_bias


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


def fast_gelu(x):
    """tanh approximation written out (reference ``training/activations.py``)."""
    return 0.5 * x * (1.0 + torch.tanh(x * 0.7978845608 * (1.0 + 0.044715 * x * x)))
