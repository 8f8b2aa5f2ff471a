"""Sparse upcycling: start a Mixture-of-Experts model from a trained dense checkpoint.

Capability parity with reference ``transformer/moe/upcycling_utils.py`` (``upcycle_state_dict`` :137,
``load_and_upcycle_model`` :166, flag ``--moe-use-upcycling``): every expert of a layer starts as a copy of the dense
MLP of that layer, the router keeps its fresh initialisation, everything else (attention, norms, embeddings) is taken
over unchanged.  Works for both expert containers of this framework -- ``SequentialMLP`` (one module per expert) and
``GroupedMLP`` (stacked ``weight1`` / ``weight2``) -- and under expert parallelism (each rank fills its local experts)
and expert tensor parallelism (the dense weights are already TP-sharded the same way).

With ``moe_ffn_hidden_size < ffn_hidden_size`` ("granular" upcycling) the dense FFN is cut into ``ffn / moe_ffn`` slices
and expert ``e`` receives slice ``e % n_slices`` (gate and up halves of a gated fc1 are sliced separately).
"""
from __future__ import annotations

import re
from typing import Dict, List

import torch

from ..parallel import state as ps

_FC = re.compile(r"^(.*\.mlp)\.linear_fc([12])\.(weight|bias)$")


def _unwrap(m):
    while hasattr(m, "module"):
        m = m.module
    return m


def _slice_fc1(w: torch.Tensor, gated: bool, n_slices: int, idx: int) -> torch.Tensor:
    if n_slices == 1:
        return w
    if gated:
        g, u = w.chunk(2, dim=0)
        return torch.cat([g.chunk(n_slices, dim=0)[idx], u.chunk(n_slices, dim=0)[idx]], dim=0)
    return w.chunk(n_slices, dim=0)[idx]


def _convert_to_moe_state_dict(dense_sd: Dict[str, torch.Tensor], moe_model) -> Dict[str, torch.Tensor]:
    moe = _unwrap(moe_model)
    cfg = moe.config
    moe_sd = moe.state_dict()
    out: Dict[str, torch.Tensor] = {}
    gated = bool(cfg.gated_linear_unit)
    for key, val in dense_sd.items():
        m = _FC.match(key)
        if m is None:
            if key in moe_sd:
                out[key] = val
            continue
        prefix, which, kind = m.group(1), m.group(2), m.group(3)
        router_key = f"{prefix}.router.weight"
        if router_key not in moe_sd:                      # this layer stays dense in the MoE model (moe_layer_freq)
            out[key] = val
            continue
        out[router_key] = moe_sd[router_key].detach().clone()
        # local experts of this rank
        seq_keys = sorted(k for k in moe_sd if k.startswith(f"{prefix}.experts.local_experts.") and
                          k.endswith(f"linear_fc{which}.{kind}"))
        if seq_keys:
            tgt_rows = moe_sd[seq_keys[0]].shape
            n_slices = max(1, (val.shape[0] if which == "1" else val.shape[-1]) // (tgt_rows[0] if which == "1" else tgt_rows[-1]))
            for i, k in enumerate(seq_keys):
                e_global = ps.get_expert_model_parallel_rank() * len(seq_keys) + i
                if which == "1":
                    out[k] = _slice_fc1(val, gated, n_slices, e_global % n_slices).clone()
                elif kind == "weight":
                    out[k] = (val.chunk(n_slices, dim=1)[e_global % n_slices] if n_slices > 1 else val).clone()
                else:
                    out[k] = val.clone()
            continue
        grouped_key = f"{prefix}.experts.weight{which}"
        if grouped_key in moe_sd and kind == "weight":
            stacked = moe_sd[grouped_key]
            n_local = moe.decoder.layers[0].mlp.num_local_experts if hasattr(moe, "decoder") else None
            per = stacked.shape[0] // (n_local or 1)
            n_slices = max(1, (val.shape[0] // per) if which == "1" else (val.shape[1] // stacked.shape[1]))
            parts = []
            for i in range(stacked.shape[0] // per):
                e_global = ps.get_expert_model_parallel_rank() * (stacked.shape[0] // per) + i
                if which == "1":
                    parts.append(_slice_fc1(val, gated, n_slices, e_global % n_slices))
                else:
                    parts.append(val.chunk(n_slices, dim=1)[e_global % n_slices] if n_slices > 1 else val)
            out[grouped_key] = torch.cat(parts, dim=0).clone()
        elif grouped_key in moe_sd:
            continue                                       # GroupedMLP has no biases
        else:
            raise KeyError(f"cannot place dense parameter {key} in the MoE model")
    for k, v in moe_sd.items():                            # anything MoE-only that the dense model does not have
        out.setdefault(k, v.detach().clone())
    return out


def upcycle_state_dict(moe_model: List, dense_model: List) -> Dict[str, Dict[str, torch.Tensor]]:
    if len(moe_model) != len(dense_model):
        raise ValueError("dense and MoE models must have the same number of (virtual pipeline) chunks")
    sd = {}
    for i, (m, d) in enumerate(zip(moe_model, dense_model)):
        if len(moe_model) > 1:
            ps.set_virtual_pipeline_model_parallel_rank(i)
        sd["model" if len(moe_model) == 1 else f"model{i}"] = _convert_to_moe_state_dict(_unwrap(d).state_dict(), m)
    return sd


def load_and_upcycle_model(load_dense_ckpt_func, moe_model: List, dense_model: List, strict: bool = True, load_args=(),
                           load_kwargs=None):
    """Load the dense checkpoint into ``dense_model`` with ``load_dense_ckpt_func`` and seed ``moe_model`` from it."""
    iteration, flops = load_dense_ckpt_func(*load_args, **(load_kwargs or {}))
    sd = upcycle_state_dict(moe_model, dense_model)
    for i, m in enumerate(moe_model):
        if len(moe_model) > 1:
            ps.set_virtual_pipeline_model_parallel_rank(i)
        _unwrap(m).load_state_dict(sd["model" if len(moe_model) == 1 else f"model{i}"], strict=strict)
    return iteration, flops
