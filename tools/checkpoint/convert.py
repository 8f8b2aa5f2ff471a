"""Checkpoint format / parallelism converter.

    python tools/checkpoint/convert.py --loader core --saver core --load-dir A --save-dir B \
        --target-tensor-parallel-size 4 --target-pipeline-parallel-size 2
    python tools/checkpoint/convert.py --loader hf_llama --saver core --load-dir <hf dir> --save-dir B ...
    python tools/checkpoint/convert.py --loader core --saver hf_llama --load-dir A --save-dir <hf dir>

Parity: reference ``tools/checkpoint/convert.py`` + ``loader_{mcore,llama_mistral,...}.py`` /
``saver_{mcore,...}.py`` (plugin pairs exchanging a canonical, unsharded description of the
model) and ``test_scripts/gpt2_convert.py`` (HF GPT-2 -> Megatron).  Here the interchange object is
a plain dict ``{"config": {...}, "weights": {canonical name: full tensor}}`` using this
framework's parameter names with global layer indices; loaders merge TP/PP shards into it and
savers re-shard it.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import sys
from types import SimpleNamespace
from typing import Dict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

COLUMN = ("linear_qkv.weight", "linear_qkv.bias", "linear_fc1.weight", "linear_fc1.bias", "word_embeddings.weight",
          "output_layer.weight", "linear_q.weight", "linear_kv.weight")
ROW = ("linear_proj.weight", "linear_fc2.weight")


def _kind(name: str) -> str:
    if name.endswith(COLUMN):
        return "col"
    if name.endswith(ROW):
        return "row"
    return "rep"


# ------------------------------------------------------------------------------- TP split / merge
def split_tp(name: str, w: torch.Tensor, tp: int, glu: bool):
    k = _kind(name)
    if tp == 1 or k == "rep":
        return [w.clone() for _ in range(tp)]
    if k == "row":
        return [c.contiguous() for c in torch.chunk(w, tp, dim=1)]
    if glu and "linear_fc1" in name:
        gate, up = torch.chunk(w, 2, dim=0)
        return [torch.cat([g, u], dim=0).contiguous() for g, u in zip(torch.chunk(gate, tp, 0), torch.chunk(up, tp, 0))]
    return [c.contiguous() for c in torch.chunk(w, tp, dim=0)]


def merge_tp(name: str, shards, glu: bool):
    k = _kind(name)
    if len(shards) == 1 or k == "rep":
        return shards[0]
    if k == "row":
        return torch.cat(shards, dim=1)
    if glu and "linear_fc1" in name:
        gates, ups = zip(*(torch.chunk(s, 2, dim=0) for s in shards))
        return torch.cat(list(gates) + list(ups), dim=0)
    return torch.cat(shards, dim=0)


_LAYER = re.compile(r"^(.*decoder\.layers\.)(\d+)(\..*)$")
_EXPERT = re.compile(r"^(.*\.experts\.local_experts\.)(\d+)(\..*)$")


def layers_per_stage(num_layers: int, pp: int):
    assert num_layers % pp == 0, "num_layers must divide by the pipeline size"
    return num_layers // pp


# ------------------------------------------------------------------------------------- core format
def load_core(load_dir: str, iteration=None) -> Dict:
    from megatronapp_b200.training import checkpointing as ck

    if iteration is None:
        with open(ck.get_checkpoint_tracker_filename(load_dir)) as f:
            txt = f.read().strip()
        iteration = 0 if txt == "release" else int(txt)
    base = ck.get_checkpoint_name(load_dir, iteration, release=(iteration == 0), return_base_dir=True)
    dirs = sorted(d for d in os.listdir(base) if d.startswith("mp_rank_"))
    first = torch.load(os.path.join(base, dirs[0], "model_optim_rng.pt"), map_location="cpu", weights_only=False)
    a0 = first.get("args")
    ep = int(getattr(a0, "expert_model_parallel_size", 1) or 1)
    tp = 1 + max(int(d.split("_")[2]) for d in dirs)
    n_extra = max(len(d.split("_")) - 3 for d in dirs)            # trailing fields: [pp rank][ep rank]
    has_pp = n_extra == 2 or (n_extra == 1 and ep == 1)
    pp = 1 + max((int(d.split("_")[3]) for d in dirs), default=0) if has_pp else 1
    shards = {}
    args = None
    for d in dirs:
        parts = d.split("_")
        t = int(parts[2])
        p = int(parts[3]) if has_pp else 0
        e = int(parts[-1]) if ep > 1 else 0
        sd = first if d == dirs[0] else torch.load(os.path.join(base, d, "model_optim_rng.pt"), map_location="cpu", weights_only=False)
        model_sd = sd["model"]
        if isinstance(model_sd.get("language_model"), dict):       # legacy (pre-mcore) nested naming
            from megatronapp_b200.legacy.state_dict import legacy_to_core

            model_sd = legacy_to_core(model_sd)
        n_local = 1 + max((int(m.group(2)) for m in map(_EXPERT.match, model_sd) if m), default=-1)
        glob = {}
        for k, v in model_sd.items():                              # local expert index -> global
            m = _EXPERT.match(k)
            glob[f"{m.group(1)}{int(m.group(2)) + e * n_local}{m.group(3)}" if m else k] = v
        if e == 0:
            shards[(t, p)] = glob
        else:
            shards[(t, p)].update({k: v for k, v in glob.items() if _EXPERT.match(k)})
        args = args or sd.get("args")
    cfg = _config_from_args(args)
    glu = cfg["gated_linear_unit"]
    lps = layers_per_stage(cfg["num_layers"], pp)
    weights = {}
    for p in range(pp):
        for name in shards[(0, p)]:
            if name.endswith("_extra_state"):
                continue
            full = merge_tp(name, [shards[(t, p)][name] for t in range(tp)], glu)
            m = _LAYER.match(name)
            gname = f"{m.group(1)}{int(m.group(2)) + p * lps}{m.group(3)}" if m else name
            if gname in weights:      # tied embedding copy on the last stage
                continue
            weights[gname] = full
    return {"config": cfg, "weights": weights, "iteration": iteration, "args": args}


def _config_from_args(args) -> Dict:
    g = lambda k, d=None: getattr(args, k, d)
    return {"num_layers": g("num_layers"), "hidden_size": g("hidden_size"), "num_attention_heads": g("num_attention_heads"),
            "num_query_groups": g("num_query_groups") or g("num_attention_heads"), "ffn_hidden_size": g("ffn_hidden_size"),
            "kv_channels": g("kv_channels") or g("hidden_size") // g("num_attention_heads"),
            "gated_linear_unit": bool(g("swiglu", False)), "normalization": g("normalization", "LayerNorm"),
            "add_bias_linear": not g("disable_bias_linear", False), "padded_vocab_size": g("padded_vocab_size"),
            "max_position_embeddings": g("max_position_embeddings"), "position_embedding_type": g("position_embedding_type", "learned_absolute"),
            "untie_embeddings_and_output_weights": bool(g("untie_embeddings_and_output_weights", False)),
            "rotary_base": g("rotary_base", 10000), "layernorm_epsilon": g("norm_epsilon", g("layernorm_epsilon", 1e-5)),
            "num_experts": g("num_experts"), "moe_router_topk": g("moe_router_topk", 2),
            "moe_ffn_hidden_size": g("moe_ffn_hidden_size") or g("ffn_hidden_size")}


def save_core(ckpt: Dict, save_dir: str, tp: int, pp: int, legacy: bool = False, ep: int = 1):
    from megatronapp_b200.training import checkpointing as ck

    cfg, weights = ckpt["config"], ckpt["weights"]
    glu = cfg["gated_linear_unit"]
    lps = layers_per_stage(cfg["num_layers"], pp)
    it = ckpt.get("iteration", 0)
    args = ckpt.get("args") or SimpleNamespace()
    args = SimpleNamespace(**{**vars(args), **_args_from_config(cfg), "tensor_model_parallel_size": tp,
                              "pipeline_model_parallel_size": pp, "expert_model_parallel_size": ep})
    n_exp = cfg.get("num_experts") or 0
    assert ep == 1 or (n_exp and n_exp % ep == 0), f"{n_exp} experts do not divide over expert-parallel size {ep}"
    n_local = n_exp // ep if n_exp else 0
    vocab = weights["embedding.word_embeddings.weight"].shape[0]
    assert vocab % tp == 0, f"vocab {vocab} must divide by target TP {tp} (re-pad first)"
    for p in range(pp):
        stage = {}
        for name, w in weights.items():
            m = _LAYER.match(name)
            if m:
                li = int(m.group(2))
                if li // lps != p:
                    continue
                lname = f"{m.group(1)}{li - p * lps}{m.group(3)}"
            elif name.startswith("embedding."):
                if p != 0:
                    continue
                lname = name
            else:      # final norm / output layer live on the last stage
                if p != pp - 1:
                    continue
                lname = name
            stage[lname] = split_tp(name, w, tp, glu)
        for t, e in ((t, e) for t in range(tp) for e in range(ep)):
            path = ck.get_checkpoint_name(save_dir, it, release=(it == 0), pipeline_parallel=pp > 1, tensor_rank=t,
                                          pipeline_rank=p, expert_parallel=ep > 1, expert_rank=e)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            model_sd = {}
            for k, v in stage.items():                   # dense tensors on every EP rank, experts on their owner only
                m = _EXPERT.match(k)
                if m is None:
                    model_sd[k] = v[t]
                elif ep == 1:
                    model_sd[k] = v[t]
                elif int(m.group(2)) // n_local == e:
                    model_sd[f"{m.group(1)}{int(m.group(2)) - e * n_local}{m.group(3)}"] = v[t]
            if legacy:
                from megatronapp_b200.legacy.state_dict import core_to_legacy

                model_sd = core_to_legacy(model_sd)
            torch.save({"model": model_sd, "args": args, "iteration": it,
                        "checkpoint_version": 3.0}, path)
    with open(ck.get_checkpoint_tracker_filename(save_dir), "w") as f:
        f.write(str(it) if it else "release")


def _args_from_config(cfg):
    return {"num_layers": cfg["num_layers"], "hidden_size": cfg["hidden_size"], "num_attention_heads": cfg["num_attention_heads"],
            "num_query_groups": cfg["num_query_groups"], "group_query_attention": cfg["num_query_groups"] != cfg["num_attention_heads"],
            "ffn_hidden_size": cfg["ffn_hidden_size"], "kv_channels": cfg["kv_channels"], "swiglu": cfg["gated_linear_unit"],
            "normalization": cfg["normalization"], "disable_bias_linear": not cfg["add_bias_linear"],
            "padded_vocab_size": cfg["padded_vocab_size"], "max_position_embeddings": cfg["max_position_embeddings"],
            "position_embedding_type": cfg["position_embedding_type"],
            "untie_embeddings_and_output_weights": cfg["untie_embeddings_and_output_weights"],
            "rotary_base": cfg["rotary_base"], "norm_epsilon": cfg["layernorm_epsilon"],
            "num_experts": cfg.get("num_experts"), "moe_router_topk": cfg.get("moe_router_topk", 2),
            "moe_ffn_hidden_size": cfg.get("moe_ffn_hidden_size")}


# ------------------------------------------------------------------------------------ HF: Llama
def _hf_state(load_dir):
    sd = {}
    files = sorted(f for f in os.listdir(load_dir) if f.endswith((".safetensors", ".bin", ".pt")))
    for f in files:
        p = os.path.join(load_dir, f)
        if f.endswith(".safetensors"):
            from safetensors.torch import load_file

            sd.update(load_file(p))
        else:
            sd.update(torch.load(p, map_location="cpu", weights_only=True))
    return sd


def qkv_from_separate(q, k, v, nh, ng, hn):
    """[nh*hn, h], [ng*hn, h] x2 -> Megatron's per-group interleaved [ng * (nh/ng + 2) * hn, h]."""
    qpg = nh // ng
    q = q.view(ng, qpg * hn, -1)
    k = k.view(ng, hn, -1)
    v = v.view(ng, hn, -1)
    return torch.cat([q, k, v], dim=1).reshape(ng * (qpg + 2) * hn, -1)


def qkv_to_separate(w, nh, ng, hn):
    qpg = nh // ng
    w = w.view(ng, (qpg + 2) * hn, -1)
    q, k, v = torch.split(w, [qpg * hn, hn, hn], dim=1)
    return q.reshape(nh * hn, -1), k.reshape(ng * hn, -1), v.reshape(ng * hn, -1)


def _hf_llama_trunk(load_dir: str):
    """Everything Llama / Mistral / Mixtral share: embeddings, norms, attention.  Returns (hf config, hf state, weights, cfg)."""
    hc = json.load(open(os.path.join(load_dir, "config.json")))
    sd = _hf_state(load_dir)
    nh, ng = hc["num_attention_heads"], hc.get("num_key_value_heads", hc["num_attention_heads"])
    h, L = hc["hidden_size"], hc["num_hidden_layers"]
    hn = hc.get("head_dim") or h // nh
    w = {"embedding.word_embeddings.weight": sd["model.embed_tokens.weight"],
         "decoder.final_layernorm.weight": sd["model.norm.weight"]}
    tied = hc.get("tie_word_embeddings", False) or "lm_head.weight" not in sd
    if not tied:
        w["output_layer.weight"] = sd["lm_head.weight"]
    for i in range(L):
        a, b = f"model.layers.{i}.", f"decoder.layers.{i}."
        w[b + "input_layernorm.weight"] = sd[a + "input_layernorm.weight"]
        w[b + "self_attention.linear_qkv.weight"] = qkv_from_separate(
            sd[a + "self_attn.q_proj.weight"], sd[a + "self_attn.k_proj.weight"], sd[a + "self_attn.v_proj.weight"], nh, ng, hn)
        w[b + "self_attention.linear_proj.weight"] = sd[a + "self_attn.o_proj.weight"]
        w[b + "pre_mlp_layernorm.weight"] = sd[a + "post_attention_layernorm.weight"]
    cfg = {"num_layers": L, "hidden_size": h, "num_attention_heads": nh, "num_query_groups": ng,
           "ffn_hidden_size": hc["intermediate_size"], "kv_channels": hn, "gated_linear_unit": True,
           "normalization": "RMSNorm", "add_bias_linear": False, "padded_vocab_size": hc["vocab_size"],
           "max_position_embeddings": hc.get("max_position_embeddings", 4096), "position_embedding_type": "rope",
           "untie_embeddings_and_output_weights": not tied, "rotary_base": hc.get("rope_theta") or (hc.get("rope_parameters") or {}).get("rope_theta", 10000.0),
           "layernorm_epsilon": hc.get("rms_norm_eps", 1e-5)}
    return hc, sd, w, cfg


def load_hf_llama(load_dir: str) -> Dict:
    hc, sd, w, cfg = _hf_llama_trunk(load_dir)
    for i in range(cfg["num_layers"]):
        a, b = f"model.layers.{i}.", f"decoder.layers.{i}."
        w[b + "mlp.linear_fc1.weight"] = torch.cat([sd[a + "mlp.gate_proj.weight"], sd[a + "mlp.up_proj.weight"]], dim=0)
        w[b + "mlp.linear_fc2.weight"] = sd[a + "mlp.down_proj.weight"]
    return {"config": cfg, "weights": w, "iteration": 0}


def load_hf_mixtral(load_dir: str) -> Dict:
    """HF Mixtral (reference ``loader_mixtral_hf.py``): the Llama trunk + ``block_sparse_moe`` (router ``gate``; per
    expert ``w1`` = gate projection, ``w3`` = up projection, ``w2`` = down projection)."""
    hc, sd, w, cfg = _hf_llama_trunk(load_dir)
    E = hc["num_local_experts"]
    fused = "model.layers.0.mlp.experts.gate_up_proj" in sd      # transformers >= 5 in-memory layout (stacked experts)
    for i in range(cfg["num_layers"]):
        b = f"decoder.layers.{i}.mlp."
        if fused:
            a = f"model.layers.{i}.mlp."
            w[b + "router.weight"] = sd[a + "gate.weight"]
            for e in range(E):
                w[f"{b}experts.local_experts.{e}.linear_fc1.weight"] = sd[a + "experts.gate_up_proj"][e]     # [gate ; up]
                w[f"{b}experts.local_experts.{e}.linear_fc2.weight"] = sd[a + "experts.down_proj"][e]
            continue
        a = f"model.layers.{i}.block_sparse_moe."                # hub checkpoint layout
        w[b + "router.weight"] = sd[a + "gate.weight"]
        for e in range(E):
            w[f"{b}experts.local_experts.{e}.linear_fc1.weight"] = torch.cat([sd[f"{a}experts.{e}.w1.weight"],
                                                                             sd[f"{a}experts.{e}.w3.weight"]], dim=0)
            w[f"{b}experts.local_experts.{e}.linear_fc2.weight"] = sd[f"{a}experts.{e}.w2.weight"]
    cfg.update(num_experts=E, moe_router_topk=hc.get("num_experts_per_tok", 2), moe_ffn_hidden_size=hc["intermediate_size"])
    return {"config": cfg, "weights": w, "iteration": 0}


def _hf_llama_trunk_out(ckpt: Dict):
    cfg, w = ckpt["config"], ckpt["weights"]
    nh, ng, hn = cfg["num_attention_heads"], cfg["num_query_groups"], cfg["kv_channels"]
    sd = {"model.embed_tokens.weight": w["embedding.word_embeddings.weight"], "model.norm.weight": w["decoder.final_layernorm.weight"]}
    if "output_layer.weight" in w:
        sd["lm_head.weight"] = w["output_layer.weight"]
    for i in range(cfg["num_layers"]):
        a, b = f"model.layers.{i}.", f"decoder.layers.{i}."
        q, k, v = qkv_to_separate(w[b + "self_attention.linear_qkv.weight"], nh, ng, hn)
        sd[a + "self_attn.q_proj.weight"], sd[a + "self_attn.k_proj.weight"], sd[a + "self_attn.v_proj.weight"] = q, k, v
        sd[a + "self_attn.o_proj.weight"] = w[b + "self_attention.linear_proj.weight"]
        sd[a + "input_layernorm.weight"] = w[b + "input_layernorm.weight"]
        sd[a + "post_attention_layernorm.weight"] = w[b + "pre_mlp_layernorm.weight"]
    hc = {"architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_size": cfg["hidden_size"],
          "num_hidden_layers": cfg["num_layers"], "num_attention_heads": nh, "num_key_value_heads": ng,
          "intermediate_size": cfg["ffn_hidden_size"], "vocab_size": cfg["padded_vocab_size"], "head_dim": hn,
          "max_position_embeddings": cfg["max_position_embeddings"], "rope_theta": cfg["rotary_base"],
          "rms_norm_eps": cfg["layernorm_epsilon"], "tie_word_embeddings": "output_layer.weight" not in w,
          "hidden_act": "silu", "torch_dtype": "bfloat16"}
    return sd, hc


def _write_hf(sd: Dict, hc: Dict, save_dir: str):
    os.makedirs(save_dir, exist_ok=True)
    torch.save({k: v.contiguous() for k, v in sd.items()}, os.path.join(save_dir, "pytorch_model.bin"))
    json.dump(hc, open(os.path.join(save_dir, "config.json"), "w"), indent=1)


def save_hf_llama(ckpt: Dict, save_dir: str):
    sd, hc = _hf_llama_trunk_out(ckpt)
    w = ckpt["weights"]
    for i in range(ckpt["config"]["num_layers"]):
        a, b = f"model.layers.{i}.", f"decoder.layers.{i}."
        gate, up = torch.chunk(w[b + "mlp.linear_fc1.weight"], 2, dim=0)
        sd[a + "mlp.gate_proj.weight"], sd[a + "mlp.up_proj.weight"] = gate, up
        sd[a + "mlp.down_proj.weight"] = w[b + "mlp.linear_fc2.weight"]
    _write_hf(sd, hc, save_dir)


def save_hf_mixtral(ckpt: Dict, save_dir: str):
    sd, hc = _hf_llama_trunk_out(ckpt)
    cfg, w = ckpt["config"], ckpt["weights"]
    E = cfg["num_experts"]
    for i in range(cfg["num_layers"]):
        a, b = f"model.layers.{i}.block_sparse_moe.", f"decoder.layers.{i}.mlp."
        sd[a + "gate.weight"] = w[b + "router.weight"]
        for e in range(E):
            g, u = torch.chunk(w[f"{b}experts.local_experts.{e}.linear_fc1.weight"], 2, dim=0)
            sd[f"{a}experts.{e}.w1.weight"], sd[f"{a}experts.{e}.w3.weight"] = g, u
            sd[f"{a}experts.{e}.w2.weight"] = w[f"{b}experts.local_experts.{e}.linear_fc2.weight"]
    hc.update({"architectures": ["MixtralForCausalLM"], "model_type": "mixtral", "num_local_experts": E,
               "num_experts_per_tok": cfg.get("moe_router_topk", 2), "intermediate_size": cfg["moe_ffn_hidden_size"]})
    _write_hf(sd, hc, save_dir)


# ------------------------------------------------------------------------------------ HF: LLaVA
# reference tools/checkpoint/{loader_llava,saver_llava,saver_hf_llava}.py: CLIP vision tower + 2-layer MLP projector +
# Llama language model.  Core-side keys: ``vision_model.*`` (CLIPViTModel), ``vision_projection.encoder.*``,
# ``language_model.*`` (GPTModel).  Both HF prefix generations are accepted (``model.vision_tower...`` and the older
# ``vision_tower...`` / ``language_model.model...``).
_LLAVA_VIT = (("input_layernorm", "layer_norm1"), ("pre_mlp_layernorm", "layer_norm2"), ("self_attention.linear_proj", "self_attn.out_proj"),
              ("mlp.linear_fc1", "mlp.fc1"), ("mlp.linear_fc2", "mlp.fc2"))


def _strip_llava_prefixes(sd: Dict) -> Dict:
    out = {}
    for k, v in sd.items():
        k2 = k[len("model."):] if k.startswith("model.") and not k.startswith("model.layers") else k
        k2 = k2.replace("language_model.model.", "language_model.").replace("language_model.lm_head.", "lm_head.")
        out[k2] = v
    return out


def load_hf_llava(load_dir: str) -> Dict:
    hc = json.load(open(os.path.join(load_dir, "config.json")))
    vc, tc = hc["vision_config"], hc["text_config"]
    sd = _strip_llava_prefixes(_hf_state(load_dir))
    w = {}
    # ---- vision tower (CLIP ViT: MHA with biases, learned positions, class token, pre LayerNorm)
    nh, h = vc["num_attention_heads"], vc["hidden_size"]
    hn = h // nh
    vt = "vision_tower.vision_model."
    w["vision_model.conv1.weight"] = sd[vt + "embeddings.patch_embedding.weight"]
    w["vision_model.class_token"] = sd[vt + "embeddings.class_embedding"].reshape(1, 1, h)
    w["vision_model.position_embeddings.weight"] = sd[vt + "embeddings.position_embedding.weight"]
    for nm in ("weight", "bias"):
        w[f"vision_model.ln_pre.{nm}"] = sd[vt + f"pre_layrnorm.{nm}"]
    for i in range(vc["num_hidden_layers"]):
        a, b = f"{vt}encoder.layers.{i}.", f"vision_model.decoder.layers.{i}."
        for nm in ("weight", "bias"):
            q, k, v = (sd[a + f"self_attn.{p}_proj.{nm}"] for p in "qkv")
            if nm == "bias":
                q, k, v = (t.unsqueeze(-1) for t in (q, k, v))
            fused = qkv_from_separate(q, k, v, nh, nh, hn)
            w[b + f"self_attention.linear_qkv.{nm}"] = fused.squeeze(-1) if nm == "bias" else fused
            for core, hf in _LLAVA_VIT:
                w[b + f"{core}.{nm}"] = sd[a + f"{hf}.{nm}"]
    # ---- projector
    for j, name in ((1, "linear_fc1"), (2, "linear_fc2")):
        for nm in ("weight", "bias"):
            w[f"vision_projection.encoder.{name}.{nm}"] = sd[f"multi_modal_projector.linear_{j}.{nm}"]
    # ---- language model: reuse the Llama mapping on the language sub-dict
    lsd = {("model." + k[len("language_model."):]): v for k, v in sd.items() if k.startswith("language_model.")}
    if "lm_head.weight" in sd:
        lsd["lm_head.weight"] = sd["lm_head.weight"]
    tmp = os.path.join(load_dir, "._llava_text_tmp")
    os.makedirs(tmp, exist_ok=True)
    try:
        torch.save(lsd, os.path.join(tmp, "pytorch_model.bin"))
        json.dump(tc, open(os.path.join(tmp, "config.json"), "w"))
        text = load_hf_llama(tmp)
    finally:
        import shutil

        shutil.rmtree(tmp, ignore_errors=True)
    for k, v in text["weights"].items():
        w["language_model." + k] = v
    cfg = dict(text["config"])
    cfg.update(model_kind="llava", vision=dict(num_layers=vc["num_hidden_layers"], hidden_size=h, num_attention_heads=nh,
                                                ffn_hidden_size=vc["intermediate_size"], patch_dim=vc.get("patch_size", 14),
                                                img_h=vc.get("image_size", 336), img_w=vc.get("image_size", 336)),
               image_token_index=hc.get("image_token_index", hc.get("image_token_id", -200)))
    return {"config": cfg, "weights": w, "iteration": 0}


def save_hf_llava(ckpt: Dict, save_dir: str):
    cfg, w = ckpt["config"], ckpt["weights"]
    vis = cfg["vision"]
    nh, h = vis["num_attention_heads"], vis["hidden_size"]
    hn = h // nh
    text = {"config": {k: v for k, v in cfg.items() if k not in ("vision", "model_kind", "image_token_index")},
            "weights": {k[len("language_model."):]: v for k, v in w.items() if k.startswith("language_model.")}}
    lsd, tc = _hf_llama_trunk_out(text)
    for i in range(text["config"]["num_layers"]):
        a, b = f"model.layers.{i}.", f"decoder.layers.{i}."
        g, u = torch.chunk(text["weights"][b + "mlp.linear_fc1.weight"], 2, dim=0)
        lsd[a + "mlp.gate_proj.weight"], lsd[a + "mlp.up_proj.weight"] = g, u
        lsd[a + "mlp.down_proj.weight"] = text["weights"][b + "mlp.linear_fc2.weight"]
    sd = {}
    for k, v in lsd.items():
        sd["lm_head.weight" if k == "lm_head.weight" else "model.language_model." + k[len("model."):]] = v
    vt = "model.vision_tower.vision_model."
    sd[vt + "embeddings.patch_embedding.weight"] = w["vision_model.conv1.weight"]
    sd[vt + "embeddings.class_embedding"] = w["vision_model.class_token"].reshape(h)
    sd[vt + "embeddings.position_embedding.weight"] = w["vision_model.position_embeddings.weight"]
    for nm in ("weight", "bias"):
        sd[vt + f"pre_layrnorm.{nm}"] = w[f"vision_model.ln_pre.{nm}"]
    for i in range(vis["num_layers"]):
        a, b = f"{vt}encoder.layers.{i}.", f"vision_model.decoder.layers.{i}."
        for nm in ("weight", "bias"):
            fused = w[b + f"self_attention.linear_qkv.{nm}"]
            q, k, v = qkv_to_separate(fused.unsqueeze(-1) if nm == "bias" else fused, nh, nh, hn)
            for p, t in zip("qkv", (q, k, v)):
                sd[a + f"self_attn.{p}_proj.{nm}"] = t.squeeze(-1) if nm == "bias" else t
            for core, hf in _LLAVA_VIT:
                sd[a + f"{hf}.{nm}"] = w[b + f"{core}.{nm}"]
    for j, name in ((1, "linear_fc1"), (2, "linear_fc2")):
        for nm in ("weight", "bias"):
            sd[f"model.multi_modal_projector.linear_{j}.{nm}"] = w[f"vision_projection.encoder.{name}.{nm}"]
    hc = {"architectures": ["LlavaForConditionalGeneration"], "model_type": "llava", "text_config": tc,
          "vision_config": {"model_type": "clip_vision_model", "hidden_size": h, "num_hidden_layers": vis["num_layers"],
                            "num_attention_heads": nh, "intermediate_size": vis["ffn_hidden_size"], "patch_size": vis["patch_dim"],
                            "image_size": vis["img_h"]},
          "image_token_index": cfg.get("image_token_index", -200), "projector_hidden_act": "gelu"}
    _write_hf(sd, hc, save_dir)


# ------------------------------------------------------------------------------------ HF: GPT-2
def load_hf_gpt2(load_dir: str) -> Dict:
    """HF GPT-2 (Conv1D weights are [in, out]; fused c_attn is [q | k | v] by columns)."""
    hc = json.load(open(os.path.join(load_dir, "config.json")))
    sd = {k.removeprefix("transformer."): v for k, v in _hf_state(load_dir).items()}
    h, L, nh = hc["n_embd"], hc["n_layer"], hc["n_head"]
    hn = h // nh
    w = {"embedding.word_embeddings.weight": sd["wte.weight"], "embedding.position_embeddings.weight": sd["wpe.weight"],
         "decoder.final_layernorm.weight": sd["ln_f.weight"], "decoder.final_layernorm.bias": sd["ln_f.bias"]}
    for i in range(L):
        a, b = f"h.{i}.", f"decoder.layers.{i}."
        for hf, mg in (("ln_1", "input_layernorm"), ("ln_2", "pre_mlp_layernorm")):
            w[b + mg + ".weight"], w[b + mg + ".bias"] = sd[a + hf + ".weight"], sd[a + hf + ".bias"]
        cw, cb = sd[a + "attn.c_attn.weight"].t(), sd[a + "attn.c_attn.bias"]     # [3h, h]
        q, k, v = torch.chunk(cw, 3, dim=0)
        w[b + "self_attention.linear_qkv.weight"] = qkv_from_separate(q, k, v, nh, nh, hn)
        qb, kb, vb = torch.chunk(cb, 3, dim=0)
        w[b + "self_attention.linear_qkv.bias"] = qkv_from_separate(qb.unsqueeze(1), kb.unsqueeze(1), vb.unsqueeze(1), nh, nh, hn).squeeze(1)
        w[b + "self_attention.linear_proj.weight"] = sd[a + "attn.c_proj.weight"].t().contiguous()
        w[b + "self_attention.linear_proj.bias"] = sd[a + "attn.c_proj.bias"]
        w[b + "mlp.linear_fc1.weight"] = sd[a + "mlp.c_fc.weight"].t().contiguous()
        w[b + "mlp.linear_fc1.bias"] = sd[a + "mlp.c_fc.bias"]
        w[b + "mlp.linear_fc2.weight"] = sd[a + "mlp.c_proj.weight"].t().contiguous()
        w[b + "mlp.linear_fc2.bias"] = sd[a + "mlp.c_proj.bias"]
    cfg = {"num_layers": L, "hidden_size": h, "num_attention_heads": nh, "num_query_groups": nh, "ffn_hidden_size": 4 * h,
           "kv_channels": hn, "gated_linear_unit": False, "normalization": "LayerNorm", "add_bias_linear": True,
           "padded_vocab_size": hc["vocab_size"], "max_position_embeddings": hc["n_positions"],
           "position_embedding_type": "learned_absolute", "untie_embeddings_and_output_weights": False,
           "rotary_base": 10000, "layernorm_epsilon": hc.get("layer_norm_epsilon", 1e-5)}
    return {"config": cfg, "weights": w, "iteration": 0}


def pad_vocab(ckpt: Dict, multiple: int):
    """Zero-pad the vocabulary dimension so it divides by ``multiple`` (TP x 128 by default)."""
    for k in ("embedding.word_embeddings.weight", "output_layer.weight"):
        if k in ckpt["weights"]:
            w = ckpt["weights"][k]
            v = (w.shape[0] + multiple - 1) // multiple * multiple
            if v != w.shape[0]:
                ckpt["weights"][k] = torch.cat([w, w.new_zeros(v - w.shape[0], w.shape[1])], dim=0)
                ckpt["config"]["padded_vocab_size"] = v


LOADERS = {"core": load_core, "legacy": load_core, "hf_mixtral": load_hf_mixtral, "hf_llama": load_hf_llama,
           "hf_mistral": load_hf_llama, "hf_gpt2": load_hf_gpt2, "hf_llava": load_hf_llava}


def main(argv=None):
    ap = argparse.ArgumentParser(description="checkpoint converter")
    ap.add_argument("--model-type", default="GPT", choices=["GPT"])
    ap.add_argument("--loader", default="core", choices=list(LOADERS))
    ap.add_argument("--saver", default="core", choices=["core", "legacy", "hf_llama", "hf_mixtral", "hf_llava"])
    ap.add_argument("--load-dir", required=True)
    ap.add_argument("--save-dir", required=True)
    ap.add_argument("--target-tensor-parallel-size", type=int, default=1)
    ap.add_argument("--target-pipeline-parallel-size", type=int, default=1)
    ap.add_argument("--target-expert-parallel-size", type=int, default=1)
    ap.add_argument("--make-vocab-size-divisible-by", type=int, default=128)
    ap.add_argument("--dtype", default=None, choices=[None, "bf16", "fp16", "fp32"])
    # the reference converter streams tensors through a multiprocessing queue between a loader and a saver process and compares
    # the two sides' arguments; here both halves run in one process on whole state dicts, so these are accepted and unused
    ap.add_argument("--max-queue-size", type=int, default=50, help="accepted for CLI parity (no loader / saver queue here)")
    ap.add_argument("--no-checking", action="store_false", dest="checking", help="accepted for CLI parity")
    a = ap.parse_args(argv)
    ckpt = LOADERS[a.loader](a.load_dir)
    if a.dtype:
        dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[a.dtype]
        ckpt["weights"] = {k: v.to(dt) for k, v in ckpt["weights"].items()}
    if a.saver in ("core", "legacy"):
        pad_vocab(ckpt, a.make_vocab_size_divisible_by * a.target_tensor_parallel_size)
        save_core(ckpt, a.save_dir, a.target_tensor_parallel_size, a.target_pipeline_parallel_size,
                  legacy=a.saver == "legacy", ep=a.target_expert_parallel_size)
    elif a.saver == "hf_mixtral":
        save_hf_mixtral(ckpt, a.save_dir)
    elif a.saver == "hf_llava":
        save_hf_llava(ckpt, a.save_dir)
    else:
        save_hf_llama(ckpt, a.save_dir)
    print(f"converted {a.loader} -> {a.saver}: {len(ckpt['weights'])} tensors, "
          f"tp={a.target_tensor_parallel_size} pp={a.target_pipeline_parallel_size} -> {a.save_dir}")


if __name__ == "__main__":
    main()
