"""TensorRT-LLM checkpoint export: converts this framework's GPT-family weights into the TRT-LLM
checkpoint layout (``config.json`` + ``rank{r}.safetensors``/``.pt`` per export-TP rank).

Parity: reference ``core/export/trtllm/`` (``TRTLLMHelper.get_trtllm_pretrained_config_and_model_weights``,
``single_device_trtllm_model_weights_converter.py``: layer-name map, fused-QKV regrouping from
Megatron's per-group interleave to ``[q | k | v]``, gated-MLP ``fc``/``gate`` split, TP re-sharding,
vocab padding; ``trtllm_layers.TRTLLMLayers`` naming).  Engine building is delegated to
``tensorrt_llm`` when it is importable (it is not part of this image).
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import torch


def _split_qkv(w: torch.Tensor, nh: int, ng: int, hn: int):
    qpg = nh // ng
    w = w.view(ng, (qpg + 2) * hn, -1)
    q, k, v = torch.split(w, [qpg * hn, hn, hn], dim=1)
    return q.reshape(nh * hn, -1), k.reshape(ng * hn, -1), v.reshape(ng * hn, -1)


class TRTLLMHelper:
    def __init__(self, transformer_config, model_type: str = "gpt", position_embedding_type: str = "rope",
                 max_position_embeddings: int = 4096, rotary_percentage: float = 1.0, rotary_base: int = 10000,
                 share_embeddings_and_output_weights: bool = False, activation: str = "gelu", seq_len_interpolation_factor=None):
        self.cfg = transformer_config
        self.model_type, self.position_embedding_type = model_type, position_embedding_type
        self.max_position_embeddings, self.rotary_percentage, self.rotary_base = max_position_embeddings, rotary_percentage, rotary_base
        self.share_embeddings_and_output_weights = share_embeddings_and_output_weights
        self.activation = activation

    def pretrained_config(self, vocab_size: int, dtype: str, tp: int, pp: int = 1) -> Dict:
        c = self.cfg
        arch = {"gpt": "GPTForCausalLM", "llama": "LlamaForCausalLM", "mixtral": "LlamaForCausalLM"}.get(self.model_type, "GPTForCausalLM")
        gated = bool(c.gated_linear_unit)
        return {"architecture": arch, "dtype": dtype, "logits_dtype": "float32", "vocab_size": vocab_size,
                "max_position_embeddings": self.max_position_embeddings, "hidden_size": c.hidden_size,
                "num_hidden_layers": c.num_layers, "num_attention_heads": c.num_attention_heads,
                "num_key_value_heads": c.num_query_groups, "head_size": c.kv_channels,
                "hidden_act": "swiglu" if gated and self.activation in ("silu", "swiglu") else self.activation,
                "intermediate_size": c.ffn_hidden_size, "norm_epsilon": c.layernorm_epsilon,
                "position_embedding_type": "rope_gpt_neox" if self.position_embedding_type == "rope" else "learned_absolute",
                "rotary_pct": self.rotary_percentage, "rotary_base": self.rotary_base, "bias": bool(c.add_bias_linear),
                "use_parallel_embedding": False, "share_embedding_table": self.share_embeddings_and_output_weights,
                "moe": {"num_experts": c.num_moe_experts or 0, "top_k": c.moe_router_topk if c.num_moe_experts else 0},
                "mapping": {"world_size": tp * pp, "tp_size": tp, "pp_size": pp},
                "quantization": {"quant_algo": None, "kv_cache_quant_algo": None}}

    def convert(self, state_dict: Dict[str, torch.Tensor], dtype: torch.dtype = torch.bfloat16) -> Dict[str, torch.Tensor]:
        """Full (unsharded) canonical weights -> full TRT-LLM-named weights."""
        c = self.cfg
        nh, ng, hn = c.num_attention_heads, c.num_query_groups, c.kv_channels
        out: Dict[str, torch.Tensor] = {}
        g = lambda k: state_dict[k].to(dtype)
        out["transformer.vocab_embedding.weight"] = g("embedding.word_embeddings.weight")
        if "embedding.position_embeddings.weight" in state_dict:
            out["transformer.position_embedding.weight"] = g("embedding.position_embeddings.weight")
        for i in range(c.num_layers):
            a, b = f"decoder.layers.{i}.", f"transformer.layers.{i}."
            for src, dst in (("input_layernorm", "input_layernorm"), ("pre_mlp_layernorm", "post_layernorm")):
                out[b + dst + ".weight"] = g(a + src + ".weight")
                if a + src + ".bias" in state_dict:
                    out[b + dst + ".bias"] = g(a + src + ".bias")
            q, k, v = _split_qkv(g(a + "self_attention.linear_qkv.weight"), nh, ng, hn)
            out[b + "attention.qkv.weight"] = torch.cat([q, k, v], dim=0)
            if a + "self_attention.linear_qkv.bias" in state_dict:
                qb, kb, vb = _split_qkv(g(a + "self_attention.linear_qkv.bias").unsqueeze(1), nh, ng, hn)
                out[b + "attention.qkv.bias"] = torch.cat([qb, kb, vb], dim=0).squeeze(1)
            out[b + "attention.dense.weight"] = g(a + "self_attention.linear_proj.weight")
            if a + "self_attention.linear_proj.bias" in state_dict:
                out[b + "attention.dense.bias"] = g(a + "self_attention.linear_proj.bias")
            fc1 = g(a + "mlp.linear_fc1.weight")
            if c.gated_linear_unit:
                gate, up = torch.chunk(fc1, 2, dim=0)
                out[b + "mlp.fc.weight"], out[b + "mlp.gate.weight"] = gate, up   # TRT-LLM: fc = act branch
            else:
                out[b + "mlp.fc.weight"] = fc1
                if a + "mlp.linear_fc1.bias" in state_dict:
                    out[b + "mlp.fc.bias"] = g(a + "mlp.linear_fc1.bias")
            out[b + "mlp.proj.weight"] = g(a + "mlp.linear_fc2.weight")
            if a + "mlp.linear_fc2.bias" in state_dict:
                out[b + "mlp.proj.bias"] = g(a + "mlp.linear_fc2.bias")
        out["transformer.ln_f.weight"] = g("decoder.final_layernorm.weight")
        if "decoder.final_layernorm.bias" in state_dict:
            out["transformer.ln_f.bias"] = g("decoder.final_layernorm.bias")
        out["lm_head.weight"] = g("output_layer.weight") if "output_layer.weight" in state_dict \
            else out["transformer.vocab_embedding.weight"]
        return out

    def shard(self, weights: Dict[str, torch.Tensor], tp: int, rank: int) -> Dict[str, torch.Tensor]:
        c = self.cfg
        nh, ng, hn = c.num_attention_heads, c.num_query_groups, c.kv_channels
        out = {}
        for k, w in weights.items():
            if tp == 1:
                out[k] = w
            elif k.endswith("attention.qkv.weight") or k.endswith("attention.qkv.bias"):
                q, kk, v = torch.split(w, [nh * hn, ng * hn, ng * hn], dim=0)
                out[k] = torch.cat([torch.chunk(q, tp, 0)[rank], torch.chunk(kk, tp, 0)[rank], torch.chunk(v, tp, 0)[rank]], 0)
            elif k.endswith(("mlp.fc.weight", "mlp.gate.weight", "mlp.fc.bias", "lm_head.weight")):
                out[k] = torch.chunk(w, tp, dim=0)[rank]
            elif k.endswith(("attention.dense.weight", "mlp.proj.weight")):
                out[k] = torch.chunk(w, tp, dim=1)[rank]
            else:
                out[k] = w
        return {k: v.contiguous() for k, v in out.items()}


def export_trtllm_checkpoint(helper: TRTLLMHelper, state_dict: Dict[str, torch.Tensor], out_dir: str, vocab_size: int,
                             export_tensor_parallel_size: int = 1, dtype: torch.dtype = torch.bfloat16) -> str:
    os.makedirs(out_dir, exist_ok=True)
    full = helper.convert(state_dict, dtype)
    name = {torch.bfloat16: "bfloat16", torch.float16: "float16", torch.float32: "float32"}[dtype]
    with open(os.path.join(out_dir, "config.json"), "w") as f:
        json.dump(helper.pretrained_config(vocab_size, name, export_tensor_parallel_size), f, indent=1)
    for r in range(export_tensor_parallel_size):
        shard = helper.shard(full, export_tensor_parallel_size, r)
        try:
            from safetensors.torch import save_file

            save_file(shard, os.path.join(out_dir, f"rank{r}.safetensors"))
        except Exception:
            torch.save(shard, os.path.join(out_dir, f"rank{r}.pt"))
    return out_dir


def build_engine(checkpoint_dir: str, engine_dir: str, max_batch_size: int = 8, max_input_len: int = 2048,
                 max_seq_len: int = 4096):
    try:
        import tensorrt_llm  # noqa: F401
    except ImportError as e:
        raise RuntimeError("tensorrt_llm is not installed in this environment; the exported checkpoint in "
                           f"{checkpoint_dir} can be built elsewhere with `trtllm-build --checkpoint_dir ...`") from e
    from tensorrt_llm.commands.build import main as trtllm_build

    return trtllm_build(["--checkpoint_dir", checkpoint_dir, "--output_dir", engine_dir, "--max_batch_size", str(max_batch_size),
                         "--max_input_len", str(max_input_len), "--max_seq_len", str(max_seq_len)])
