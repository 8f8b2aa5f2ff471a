"""Expert FFNs: ``SequentialMLP`` (one MLP module per local expert) and ``GroupedMLP`` (stacked
weights, one tcgen05 GEMM per expert slice of the sorted token buffer).

Parity: reference ``transformer/moe/experts.py`` (``GroupedMLP`` :90-265 with weight1
``[h, E_local * ffn(*2)]`` / weight2 ``[E_local * ffn, h]``, ``SequentialMLP`` :856).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..parallel import state as ps
from ..parallel.random import get_cuda_rng_tracker
from ..transformer.mlp import MLP, MLPSubmodules
from ..transformer.module import MegatronModule



def _expert_replica(sd: dict) -> dict:
    """Expert weights are replicated over the EXPERT data-parallel group, not the dense one: the main replica is the
    one with expert-DP rank 0."""
    import dataclasses

    edp = ps.get_expert_data_parallel_rank()
    out = {}
    for k, t in sd.items():
        rid = tuple(t.replica_id) if isinstance(t.replica_id, (tuple, list)) else (t.replica_id,)
        out[k] = dataclasses.replace(t, replica_id=rid[:-1] + (edp,), replica_count=ps.get_state().world("edp")) \
            if (len(rid) >= 1 and hasattr(t, "replica_count")) else t
    return out


def _expert_offset(num_local: int) -> int:
    return ps.get_expert_model_parallel_rank() * num_local


def _check_etp():
    if ps.get_expert_tensor_parallel_world_size() not in (1, ps.get_tensor_model_parallel_world_size()):
        raise NotImplementedError("sharded checkpoints need expert-tensor-parallel size 1 or equal to the dense TP size")


class SequentialMLP(MegatronModule):
    def __init__(self, num_local_experts, config, submodules: MLPSubmodules):
        super().__init__(config)
        self.add_bias = config.add_bias_linear
        self.num_local_experts = num_local_experts
        self.local_experts = nn.ModuleList([MLP(config, submodules, is_expert=True) for _ in range(num_local_experts)])

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        """Keys carry the GLOBAL expert index, so checkpoints reshard over the expert-parallel degree."""
        from ..transformer.module import sharded_state_dict_of

        _check_etp()
        sd, off = {}, _expert_offset(self.num_local_experts)
        for i, ex in enumerate(self.local_experts):
            sd.update(sharded_state_dict_of(ex, f"{prefix}local_experts.{off + i}.", sharded_offsets, metadata))
        return _expert_replica(sd)

    def forward(self, permuted_local_hidden_states, tokens_per_expert, permuted_probs=None):
        from .token_dispatcher import RaggedTokens

        if isinstance(permuted_local_hidden_states, RaggedTokens):
            # ragged layout of the fused dispatcher: per-expert row ranges come from the device-side offsets (this
            # module-per-expert path reads them back; GroupedMLP is the sync-free one)
            rt = permuted_local_hidden_states
            offs = rt.layout.offsets.tolist()
            out = torch.zeros(rt.data.shape[0], self.config.hidden_size, dtype=rt.data.dtype, device=rt.data.device)
            pieces = []
            for i, expert in enumerate(self.local_experts):
                r0, r1 = offs[i], offs[i + 1]
                if r1 > r0:
                    o, b = expert(rt.data[r0:r1])
                    pieces.append((r0, r1, o if b is None else o + b))
            if pieces:
                out = torch.cat([torch.zeros(0, self.config.hidden_size, dtype=out.dtype, device=out.device)] +
                                [p[2] for p in pieces] +
                                [torch.zeros(rt.data.shape[0] - pieces[-1][1], self.config.hidden_size, dtype=out.dtype,
                                             device=out.device)], dim=0)
            return RaggedTokens(out, rt.layout, rt.probs), None
        if self.num_local_experts == 1:
            out, bias = self.local_experts[0](permuted_local_hidden_states)
            return out, bias
        sizes = tokens_per_expert.tolist()
        outs = []
        for expert, chunk in zip(self.local_experts, torch.split(permuted_local_hidden_states, sizes)):
            if chunk.shape[0] == 0:
                outs.append(chunk.new_zeros(0, self.config.hidden_size))
                continue
            o, b = expert(chunk)
            if b is not None:
                o = o + b
            outs.append(o)
        return torch.cat(outs, dim=0), None


class GroupedMLP(MegatronModule):
    def __init__(self, num_local_experts, config, submodules: Optional[MLPSubmodules] = None):
        super().__init__(config)
        self.num_local_experts = num_local_experts
        if config.add_bias_linear:
            raise ValueError("bias is not supported in GroupedMLP; use --disable-bias-linear")
        etp = ps.get_expert_tensor_parallel_world_size()
        ffn = config.moe_ffn_hidden_size // etp
        self.ffn = ffn
        self.glu = config.gated_linear_unit
        fc1_out = ffn * (2 if self.glu else 1)
        dev = "cpu" if (config.use_cpu_initialization or not torch.cuda.is_available()) else torch.cuda.current_device()
        # stored per expert as [out, in] rows (K-major for the forward GEMM)
        self.weight1 = nn.Parameter(torch.empty(num_local_experts * fc1_out, config.hidden_size,
                                                dtype=config.params_dtype, device=dev))
        self.weight2 = nn.Parameter(torch.empty(num_local_experts * config.hidden_size, ffn,
                                                dtype=config.params_dtype, device=dev))
        tr = get_cuda_rng_tracker()
        ctx = tr.fork("expert-parallel-rng") if "expert-parallel-rng" in tr.states_ else torch.random.fork_rng(devices=[])
        with ctx:
            config.init_method(self.weight1)
            config.output_layer_init_method(self.weight2)
        for w, dim in ((self.weight1, 0), (self.weight2, 1)):
            w.tensor_model_parallel, w.partition_dim, w.partition_stride = True, dim, 1
            w.allreduce = not (config.expert_model_parallel_size > 1)
        self.fc1_out = fc1_out

    def sharded_state_dict(self, prefix="", sharded_offsets=(), metadata=None):
        """One entry per (global expert, matrix): the stacked parameters are exposed as per-expert views with the same
        keys ``SequentialMLP`` uses, so grouped and sequential checkpoints are interchangeable."""
        from ..ckpt.sharded import make_sharded_tensor_for_checkpoint, make_tp_sharded_tensor_for_checkpoint

        _check_etp()
        etp = ps.get_expert_tensor_parallel_world_size()
        sd, off, h = {}, _expert_offset(self.num_local_experts), self.config.hidden_size
        for i in range(self.num_local_experts):
            for name, view, axis in (("linear_fc1.weight", self.weight1.narrow(0, i * self.fc1_out, self.fc1_out), 0),
                                     ("linear_fc2.weight", self.weight2.narrow(0, i * h, h), 1)):
                key = f"{prefix}local_experts.{off + i}.{name}"
                sd[key] = (make_tp_sharded_tensor_for_checkpoint(view, key, axis, prepend_offsets=tuple(sharded_offsets))
                           if etp > 1 else make_sharded_tensor_for_checkpoint(view, key, prepend_offsets=tuple(sharded_offsets)))
        return _expert_replica(sd)

    def _activation(self, inter):
        return ops.swiglu(inter) if (self.glu and self.config.activation_func is F.silu) else (
            self._glu(inter) if self.glu else self._act(inter))

    def forward(self, permuted_local_hidden_states, tokens_per_expert, permuted_probs=None):
        from .token_dispatcher import RaggedTokens

        if isinstance(permuted_local_hidden_states, RaggedTokens):
            # ONE grouped tcgen05 GEMM per matrix over the ragged layout; the router probability rides in the FC2
            # epilogue and the un-permutation is a gather-sum: no host sync, no per-expert launches
            from ..ops import moe as mops

            rt = permuted_local_hidden_states
            fuse = bool(self.config.gradient_accumulation_fusion)
            inter = mops.grouped_linear(rt.data, self.weight1, rt.layout, self.fc1_out, fuse)
            inter = self._activation(inter)
            call = getattr(rt, "ep_call", None)
            if call is not None:      # expert parallel: outputs go home over the EP arena
                from .token_dispatcher import _GroupedLinearCombineEP

                out = _GroupedLinearCombineEP.apply(inter, self.weight2, call, self.config.hidden_size, fuse)
                return RaggedTokens(None, rt.layout, rt.probs, combined=out), None
            out = mops.grouped_linear_combine(inter, self.weight2, rt.probs, rt.layout, self.config.hidden_size, fuse)
            return RaggedTokens(None, rt.layout, rt.probs, combined=out), None
        x = permuted_local_hidden_states
        sizes = tokens_per_expert.tolist()
        h = self.config.hidden_size
        outs = []
        off = 0
        for e, n in enumerate(sizes):
            chunk = x[off:off + n]
            off += n
            if n == 0:
                outs.append(x.new_zeros(0, h))
                continue
            w1 = self.weight1[e * self.fc1_out:(e + 1) * self.fc1_out]
            w2 = self.weight2[e * h:(e + 1) * h]
            inter = ops.linear(chunk, w1)
            inter = ops.swiglu(inter) if (self.glu and self.config.activation_func is F.silu) else (
                self._glu(inter) if self.glu else self._act(inter))
            outs.append(ops.linear(inter, w2))
        return torch.cat(outs, dim=0), None

    def _act(self, t):
        return ops.bias_gelu(t) if self.config.activation_func is F.gelu else self.config.activation_func(t)

    def _glu(self, t):
        g, u = t.chunk(2, dim=-1)
        return self.config.activation_func(g) * u


TEGroupedMLP = GroupedMLP        # one grouped tcgen05 GEMM per expert matrix is the only grouped implementation here


def expert_dist_ckpt_decorator(func):
    """The reference swaps the expert-parallel groups in around ``sharded_state_dict`` calls of TE modules; the experts here build
    their ShardedTensors with explicit expert offsets, so the wrapped function is returned unchanged."""
    return func
