"""Name mapping between the legacy (``megatron/legacy/model``) nested state dict and the core flat one.

Legacy layout (reference ``legacy/model/language_model.py state_dict_for_save_checkpoint`` and
``tools/checkpoint/loader_legacy.py:283-345``)::

    {"language_model": {"embedding": {"word_embeddings": {"weight"}, "position_embeddings": {"weight"}},
                        "encoder":   {"layers.N.input_norm.weight", "layers.N.self_attention.query_key_value.weight",
                                      "layers.N.self_attention.dense.weight", "layers.N.post_attention_norm.weight",
                                      "layers.N.mlp.dense_h_to_4h.weight", "layers.N.mlp.dense_4h_to_h.weight",
                                      "final_norm.weight", ...},
                        "output_layer": {"weight"}, "pooler": {...}},
     "lm_head": {...}, "binary_head": {...}}

Tensor layouts are identical on both sides (QKV interleaved per query group, GLU halves stacked), so the mapping
is purely a rename.
"""
from __future__ import annotations

import re
from typing import Dict

import torch

_LAYER = [
    ("input_layernorm.", "input_norm."),
    ("self_attention.linear_qkv.", "self_attention.query_key_value."),
    ("self_attention.linear_proj.", "self_attention.dense."),
    ("self_attention.linear_q.", "inter_attention.query."),
    ("pre_cross_attn_layernorm.", "post_inter_attention_norm."),
    ("cross_attention.linear_q.", "inter_attention.query."),
    ("cross_attention.linear_kv.", "inter_attention.key_value."),
    ("cross_attention.linear_proj.", "inter_attention.dense."),
    ("pre_mlp_layernorm.", "post_attention_norm."),
    ("mlp.linear_fc1.", "mlp.dense_h_to_4h."),
    ("mlp.linear_fc2.", "mlp.dense_4h_to_h."),
]


def _rename_layer_key(k: str, to_legacy: bool) -> str:
    for core, leg in _LAYER:
        a, b = (core, leg) if to_legacy else (leg, core)
        if k.startswith(a):
            return b + k[len(a):]
    return k


def core_to_legacy(sd: Dict[str, torch.Tensor], stack: str = "decoder") -> Dict:
    """Flat core GPT/BERT state dict -> nested legacy dict.  ``stack`` is the core name of the transformer
    block that becomes the legacy ``encoder`` (GPT: ``decoder``; BERT: ``encoder``)."""
    lm: Dict = {"embedding": {}, "encoder": {}}
    out: Dict = {"language_model": lm}
    for k, v in sd.items():
        if k.endswith("_extra_state"):
            continue
        if k.startswith("embedding."):
            _, sub, leaf = k.split(".", 2)
            lm["embedding"].setdefault(sub, {})[leaf] = v
        elif k.startswith(f"{stack}.layers."):
            m = re.match(rf"{stack}\.layers\.(\d+)\.(.*)", k)
            lm["encoder"][f"layers.{m.group(1)}.{_rename_layer_key(m.group(2), True)}"] = v
        elif k.startswith(f"{stack}.final_layernorm."):
            lm["encoder"]["final_norm." + k.rsplit(".", 1)[1]] = v
        elif k.startswith("output_layer."):
            lm.setdefault("output_layer", {})[k.split(".", 1)[1]] = v
        elif k.startswith("pooler."):
            lm.setdefault("pooler", {})[k.split(".", 1)[1]] = v
        else:                                   # heads (lm_head.*, binary_head.*) keep their own top-level dict
            top, leaf = k.split(".", 1)
            out.setdefault(top, {})[leaf] = v
    return out


def legacy_to_core(sd: Dict, stack: str = "decoder") -> Dict[str, torch.Tensor]:
    lm = sd.get("language_model", sd)
    out: Dict[str, torch.Tensor] = {}
    for sub, leaves in lm.get("embedding", {}).items():
        for leaf, v in leaves.items():
            out[f"embedding.{sub}.{leaf}"] = v
    enc = lm.get("encoder", lm.get("transformer", {}))       # very old checkpoints call it "transformer"
    for k, v in enc.items():
        m = re.match(r"layers\.(\d+)\.(.*)", k)
        if m:
            out[f"{stack}.layers.{m.group(1)}.{_rename_layer_key(m.group(2), False)}"] = v
        elif k.startswith(("final_norm.", "final_layernorm.")):
            out[f"{stack}.final_layernorm." + k.rsplit(".", 1)[1]] = v
    for name in ("output_layer", "pooler"):
        for leaf, v in lm.get(name, {}).items():
            out[f"{name}.{leaf}"] = v
    for top, leaves in sd.items():
        if top != "language_model" and isinstance(leaves, dict):
            for leaf, v in leaves.items():
                if torch.is_tensor(v):
                    out[f"{top}.{leaf}"] = v
    return out
