"""Mamba-2 (SSD) mixer, hybrid layer allocation and the Mamba stack.

Parity: reference ``core/ssm/`` -- ``MambaMixer`` (in_proj -> causal depthwise conv1d -> selective
state-space scan with scalar-per-head A, grouped B/C, D skip, gated RMSNorm -> out_proj),
``mamba_hybrid_layer_allocation.py`` (``Symbols``: ``M`` mamba, ``*`` attention, ``-`` MLP; ratio
based allocation and explicit override patterns) and ``MambaStack``.  The scan is a chunked
(block-parallel) PyTorch implementation of the SSD recurrence; the reference calls
``mamba_ssm`` Triton kernels (SURVEY K23 marks the hand-written kernel a stretch goal).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..parallel.tp_layers import ColumnParallelLinear, RowParallelLinear
from ..transformer.module import MegatronModule
from ..transformer.norms import FusedLayerNorm
from ..transformer.spec_utils import ModuleSpec, build_module


class Symbols:
    MAMBA, ATTENTION, MLP = "M", "*", "-"
    VALID = {"M", "*", "-"}


def allocate_layers(total_layers: int, attention_ratio: float, mlp_ratio: float,
                    override_pattern: Optional[str] = None) -> List[str]:
    if override_pattern is not None:
        pat = list(override_pattern)
        if len(pat) != total_layers or not set(pat) <= Symbols.VALID:
            raise ValueError(f"hybrid override pattern {override_pattern!r} does not describe {total_layers} layers")
        return pat
    if not (0.0 <= attention_ratio <= 1.0 and 0.0 <= mlp_ratio <= 1.0 and attention_ratio + mlp_ratio <= 1.0):
        raise ValueError("attention/mlp ratios must be in [0, 1] and sum to <= 1")
    n_attn = round(total_layers * attention_ratio)
    n_mlp = round(total_layers * mlp_ratio)
    layers = [Symbols.MAMBA] * total_layers
    if n_attn:
        step = total_layers / n_attn
        for i in range(n_attn):
            layers[min(total_layers - 1, int(step * i + step / 2))] = Symbols.ATTENTION
    if n_mlp:
        free = [i for i, l in enumerate(layers) if l == Symbols.MAMBA]
        step = len(free) / n_mlp
        for i in range(n_mlp):
            layers[free[min(len(free) - 1, int(step * i + step / 2))]] = Symbols.MLP
    return layers


def ssd_scan(x, dt, A, B, C, D, chunk: int = 64, initial_state=None):
    """x [b, l, h, p], dt [b, l, h] (>0), A [h] (<0), B/C [b, l, g, n], D [h] -> y [b, l, h, p], final state.
    Recurrence per head: S_t = exp(dt_t A) S_{t-1} + dt_t x_t B_t^T ;  y_t = S_t C_t + D x_t."""
    b, l, h, p = x.shape
    g, n = B.shape[2], B.shape[3]
    rep = h // g
    Bh = B.repeat_interleave(rep, dim=2).float()
    Ch = C.repeat_interleave(rep, dim=2).float()
    xf, dtf = x.float(), dt.float()
    dA = dtf * A.float().view(1, 1, h)                               # [b, l, h] log decay per step
    state = torch.zeros(b, h, p, n, dtype=torch.float32, device=x.device) if initial_state is None else initial_state.float()
    ys = []
    for s in range(0, l, chunk):
        e = min(l, s + chunk)
        da = dA[:, s:e]                                               # [b, c, h]
        cum = torch.cumsum(da, dim=1)                                 # inclusive cumulative log decay
        xb = xf[:, s:e] * dtf[:, s:e].unsqueeze(-1)                   # dt * x
        Bc, Cc = Bh[:, s:e], Ch[:, s:e]
        # intra-chunk: y_t += sum_{u<=t} exp(cum_t - cum_u) (C_t . B_u) dt_u x_u
        L = torch.exp(cum.unsqueeze(2) - cum.unsqueeze(1))            # [b, t, u, h]
        c_len = e - s
        mask = torch.tril(torch.ones(c_len, c_len, device=x.device, dtype=torch.bool))
        L = L.masked_fill(~mask.view(1, c_len, c_len, 1), 0.0)
        G = torch.einsum("bthn,buhn->btuh", Cc, Bc) * L
        y = torch.einsum("btuh,buhp->bthp", G, xb)
        # contribution of the carried state: y_t += exp(cum_t) C_t . S_prev
        y = y + torch.einsum("bthn,bhpn->bthp", Cc * torch.exp(cum).unsqueeze(-1), state)
        # state update: S = exp(cum_end) S_prev + sum_u exp(cum_end - cum_u) dt_u x_u B_u^T
        decay_to_end = torch.exp(cum[:, -1:].expand_as(cum) - cum)   # [b, c, h]
        state = state * torch.exp(cum[:, -1]).view(b, h, 1, 1) + torch.einsum("buhp,buhn->bhpn", xb * decay_to_end.unsqueeze(-1), Bc)
        ys.append(y)
    y = torch.cat(ys, dim=1) + xf * D.float().view(1, 1, h, 1)
    return y.to(x.dtype), state


@dataclass
class MambaMixerSubmodules:
    in_proj: Union[ModuleSpec, type] = ColumnParallelLinear
    out_proj: Union[ModuleSpec, type] = RowParallelLinear


class MambaMixer(MegatronModule):
    """Mamba-2 mixer, tensor-parallel over heads (reference ``core/ssm/mamba_mixer.py``).

    Under TP every rank owns ``nheads / tp`` heads and ``ngroups / tp`` B/C groups: ``in_proj`` is column-parallel and
    its local rows are ordered ``[z | x | B | C | dt]`` *of this rank's heads*; the depth-wise conv, ``dt_bias``,
    ``A_log``, ``D`` and the gated RMS norm (normalised per group of ``d_inner / ngroups`` channels, hence local) are
    sharded the same way; ``out_proj`` is row-parallel, so one layer costs one all-gather/all-reduce pair (or, with
    sequence parallelism, the all-gather inside ``in_proj`` and the reduce-scatter inside ``out_proj``)."""

    def __init__(self, config, submodules: MambaMixerSubmodules = None, d_model=None, d_conv=4, expand=2,
                 layer_number=None, chunk_size=64):
        super().__init__(config)
        from ..parallel import state as ps
        from ..parallel.tp_layers import set_tensor_model_parallel_attributes

        submodules = submodules or MambaMixerSubmodules()
        self.d_model = d_model or config.hidden_size
        self.d_state, self.headdim, self.ngroups = config.mamba_state_dim, config.mamba_head_dim, config.mamba_num_groups
        self.d_inner = expand * self.d_model
        self.nheads = self.d_inner // self.headdim
        self.ngroups = min(self.ngroups, self.nheads)
        self.d_conv, self.chunk_size, self.layer_number = d_conv, chunk_size, layer_number
        tp = self.tp = ps.get_tensor_model_parallel_world_size()
        if self.nheads % tp or self.ngroups % tp or self.nheads % self.ngroups:
            raise ValueError(f"mamba: nheads={self.nheads} and ngroups={self.ngroups} must divide by tp={tp} "
                             f"(and nheads by ngroups)")
        self.nheads_local, self.ngroups_local, self.d_inner_local = self.nheads // tp, self.ngroups // tp, self.d_inner // tp
        conv_dim = self.d_inner_local + 2 * self.ngroups_local * self.d_state
        # z, x, B, C, dt
        self.in_proj = build_module(submodules.in_proj, self.d_model, 2 * self.d_inner + 2 * self.ngroups * self.d_state + self.nheads,
                                    config=config, init_method=config.init_method, gather_output=False, bias=False,
                                    skip_bias_add=False)
        dev, dt_ = self.in_proj.weight.device, config.params_dtype
        self.conv1d = nn.Conv1d(conv_dim, conv_dim, d_conv, groups=conv_dim, padding=d_conv - 1, bias=True, device=dev, dtype=dt_)
        self.dt_bias = nn.Parameter(torch.log(torch.expm1(torch.exp(torch.rand(self.nheads_local, device=dev) * (math.log(0.1) - math.log(0.001)) + math.log(0.001)))).to(dt_))
        self.A_log = nn.Parameter(torch.log(torch.empty(self.nheads_local, device=dev).uniform_(1, 16)).to(dt_))
        self.D = nn.Parameter(torch.ones(self.nheads_local, device=dev, dtype=dt_))
        self.norm_weight = nn.Parameter(torch.ones(self.d_inner_local, device=dev, dtype=dt_))
        self.out_proj = build_module(submodules.out_proj, self.d_inner, self.d_model, config=config,
                                     init_method=config.output_layer_init_method, bias=False, input_is_parallel=True,
                                     skip_bias_add=True)
        for p in (self.dt_bias, self.A_log, self.D):
            p._no_weight_decay = True
        for p in (self.dt_bias, self.A_log, self.D, self.norm_weight, self.conv1d.weight, self.conv1d.bias):
            set_tensor_model_parallel_attributes(p, True, 0, 1)

    def _segments(self):
        """Local row counts of the ``[z | x | B | C | dt]`` segments of ``in_proj`` and ``[x | B | C]`` of the conv."""
        gn = self.ngroups_local * self.d_state
        return ([("z", self.d_inner_local), ("x", self.d_inner_local), ("B", gn), ("C", gn), ("dt", self.nheads_local)],
                [("x", self.d_inner_local), ("B", gn), ("C", gn)])

    def forward(self, hidden_states, inference_context=None):
        zxbcdt, _ = self.in_proj(hidden_states)                         # [s, b, local]  (s is the full sequence)
        s, b, _ = zxbcdt.shape
        zxbcdt = zxbcdt.transpose(0, 1)                                # [b, s, ...]
        di, gn, nh, ng = self.d_inner_local, self.ngroups_local * self.d_state, self.nheads_local, self.ngroups_local
        z, xBC, dt = torch.split(zxbcdt, [di, di + 2 * gn, nh], dim=-1)
        ssm_state = None
        if inference_context is not None:
            # recurrent decode: carry the last (d_conv - 1) conv inputs and the SSM state between calls
            # (reference mamba_mixer.py keeps conv_state / ssm_state in the inference context the same way)
            cache = inference_context.key_value_memory_dict
            key = ("mamba", self.layer_number)
            b0 = inference_context.batch_size_offset
            if inference_context.sequence_len_offset == 0 or key not in cache:
                cache[key] = [xBC.new_zeros(inference_context.max_batch_size, self.d_conv - 1, xBC.shape[-1]),
                              torch.zeros(inference_context.max_batch_size, nh, self.headdim, self.d_state,
                                          dtype=torch.float32, device=xBC.device)]
            conv_state, ssm_all = cache[key]
            seq = torch.cat([conv_state[b0:b0 + b], xBC], dim=1)                       # [b, d_conv-1+s, c]
            conv_state[b0:b0 + b] = seq[:, -(self.d_conv - 1):]
            xBC = F.silu(F.conv1d(seq.transpose(1, 2), self.conv1d.weight, self.conv1d.bias, groups=seq.shape[-1]).transpose(1, 2))
            ssm_state = ssm_all[b0:b0 + b]
        else:
            xBC = F.silu(self.conv1d(xBC.transpose(1, 2))[..., :s].transpose(1, 2))
        x, B, C = torch.split(xBC, [di, gn, gn], dim=-1)
        dt = F.softplus(dt.float() + self.dt_bias.float())
        y, new_state = ssd_scan(x.reshape(b, s, nh, self.headdim), dt, -torch.exp(self.A_log.float()),
                                B.reshape(b, s, ng, self.d_state), C.reshape(b, s, ng, self.d_state),
                                self.D, self.chunk_size, initial_state=ssm_state)
        if inference_context is not None:
            ssm_all[b0:b0 + b] = new_state
        y = y.reshape(b, s, di)
        y = (y * F.silu(z)).float().view(b, s, ng, di // ng)            # gated RMSNorm, one statistic per B/C group
        y = (y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + 1e-5)).view(b, s, di) * self.norm_weight.float()
        out, bias = self.out_proj(y.to(hidden_states.dtype).transpose(0, 1).contiguous())
        return out, bias

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        """``in_proj`` / ``conv1d`` are stored as their logical segments (``...in_proj.weight.z`` etc.), each sharded
        over TP on its own, so a checkpoint reshards to any TP degree (reference ``mamba_mixer.py`` does the same
        with split factories)."""
        from ..ckpt.sharded import make_tp_sharded_tensor_for_checkpoint

        off = tuple(sharded_offsets)
        sd = {}
        seg_in, seg_conv = self._segments()

        def split(t, key, segs):
            o = 0
            for name, n in segs:
                sd[f"{key}.{name}"] = make_tp_sharded_tensor_for_checkpoint(t.narrow(0, o, n), f"{key}.{name}", 0,
                                                                            prepend_offsets=off)
                o += n

        split(self.in_proj.weight, f"{prefix}in_proj.weight", seg_in)
        split(self.conv1d.weight, f"{prefix}conv1d.weight", seg_conv)
        split(self.conv1d.bias, f"{prefix}conv1d.bias", seg_conv)
        for name in ("dt_bias", "A_log", "D", "norm_weight"):
            key = f"{prefix}{name}"
            sd[key] = make_tp_sharded_tensor_for_checkpoint(getattr(self, name), key, 0, prepend_offsets=off)
        sd.update(self.out_proj.sharded_state_dict(f"{prefix}out_proj.", off, metadata))
        return sd


class MambaLayer(MegatronModule):
    def __init__(self, config, layer_number=1, **_):
        super().__init__(config)
        self.norm = FusedLayerNorm(config, config.hidden_size, config.layernorm_epsilon)
        self.mixer = MambaMixer(config, layer_number=layer_number)

    def forward(self, hidden_states, attention_mask=None, inference_context=None, rotary_pos_emb=None, **_):
        out, bias = self.mixer(self.norm(hidden_states), inference_context=inference_context)
        if bias is not None:
            out = out + bias
        return hidden_states + out


class MambaStack(MegatronModule):
    def __init__(self, config, attention_layer_spec, mlp_layer_spec, pre_process=True, post_process=True,
                 hybrid_attention_ratio=0.0, hybrid_mlp_ratio=0.0, hybrid_override_pattern=None):
        super().__init__(config)
        self.pre_process, self.post_process = pre_process, post_process
        self.input_tensor = None
        from ..parallel import state as ps

        all_types = allocate_layers(config.num_layers, hybrid_attention_ratio, hybrid_mlp_ratio, hybrid_override_pattern)
        # this pipeline stage owns a contiguous slice of the layer pattern (reference MambaStack: pp_layer_offset)
        pp = ps.get_pipeline_model_parallel_world_size()
        if pp > 1:
            if ps.get_virtual_pipeline_model_parallel_world_size() is not None:
                raise ValueError("interleaved pipeline schedule is not supported for Mamba stacks")
            if config.num_layers % pp:
                raise ValueError(f"num_layers ({config.num_layers}) must divide by the pipeline size ({pp})")
        n_local = config.num_layers // pp
        self.layer_offset = ps.get_pipeline_model_parallel_rank() * n_local if pp > 1 else 0
        self.layer_types = all_types[self.layer_offset:self.layer_offset + n_local]
        layers = []
        for i, t in enumerate(self.layer_types):
            if t == Symbols.MAMBA:
                layers.append(MambaLayer(config, layer_number=self.layer_offset + i + 1))
            elif t == Symbols.ATTENTION:        # TransformerLayer adds the pipeline offset to its layer number itself
                layers.append(build_module(attention_layer_spec, config=config, layer_number=i + 1))
            else:
                layers.append(build_module(mlp_layer_spec, config=config, layer_number=i + 1))
        self.layers = nn.ModuleList(layers)
        self.final_norm = FusedLayerNorm(config, config.hidden_size, config.layernorm_epsilon) if post_process else None

    def set_input_tensor(self, t):
        self.input_tensor = t

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        from ..transformer.module import sharded_state_dict_of

        sd = {}
        for i, layer in enumerate(self.layers):            # checkpoint keys carry the GLOBAL layer index
            sd.update(sharded_state_dict_of(layer, f"{prefix}layers.{self.layer_offset + i}.", sharded_offsets, metadata))
        if self.final_norm is not None:
            sd.update(sharded_state_dict_of(self.final_norm, f"{prefix}final_norm.", sharded_offsets, metadata))
        return sd

    def forward(self, hidden_states, attention_mask=None, inference_context=None, rotary_pos_emb=None):
        if not self.pre_process:
            hidden_states = self.input_tensor
        for layer, t in zip(self.layers, self.layer_types):
            if t == Symbols.MAMBA:
                hidden_states = layer(hidden_states, inference_context=inference_context)
            else:
                hidden_states, _ = layer(hidden_states, attention_mask=attention_mask, rotary_pos_emb=rotary_pos_emb,
                                         inference_context=inference_context)
        if self.final_norm is not None:
            hidden_states = self.final_norm(hidden_states)
        return hidden_states


# -- spec dataclasses / names of the reference's ``ssm`` package ----------------------------------------------------------------------
@dataclass
class MambaLayerSubmodules:
    norm: Union[ModuleSpec, type] = FusedLayerNorm
    mixer: Union[ModuleSpec, type] = MambaMixer
    mamba_bda: Union[ModuleSpec, type, None] = None


@dataclass
class MambaStackSubmodules:
    mamba_layer: Union[ModuleSpec, type] = MambaLayer
    attention_layer: Union[ModuleSpec, type, None] = None
    mlp_layer: Union[ModuleSpec, type, None] = None


class ExtendedRMSNorm(FusedLayerNorm):
    """Gated-RMSNorm name of the reference's mixer; the grouped gated norm itself is inside ``MambaMixer.forward`` here."""

    def __init__(self, config, hidden_size: int, eps: float = 1e-5, **kw):
        import dataclasses as _dc

        super().__init__(_dc.replace(config, normalization="RMSNorm") if config is not None else None, hidden_size, eps,
                         normalization="RMSNorm")
