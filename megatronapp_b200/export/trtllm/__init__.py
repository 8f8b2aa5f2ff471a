"""TensorRT-LLM checkpoint export: converts this framework's GPT-family weights into the TRT-LLM
checkpoint layout (``config.json`` + ``rank{r}.safetensors``/``.pt`` per export-TP rank).

Parity: reference ``core/export/trtllm/`` (``TRTLLMHelper.get_trtllm_pretrained_config_and_model_weights``,
``single_device_trtllm_model_weights_converter.py``: layer-name map, fused-QKV regrouping from
Megatron's per-group interleave to ``[q | k | v]``, gated-MLP ``fc``/``gate`` split, TP re-sharding,
vocab padding; ``trtllm_layers.TRTLLMLayers`` naming; MoE expert stacking; zero-centred LayerNorm gamma; fp8 weight /
activation / KV-cache scaling factors; pipeline split of the exported layers) and
``distributed_trtllm_model_weights_converter.py`` (``DistributedTRTLLMConverter``: every training rank converts its *own*
TP / PP shard on the device -- no gather of the model -- and only the vocabulary-parallel embedding is re-assembled over
the TP group so the tokenizer padding can be removed).  The engine build plan lives in ``trtllm_engine_builder.py``.
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
        zero_centered = bool(getattr(c, "layernorm_zero_centered_gamma", False)) and getattr(c, "normalization", "LayerNorm") == "LayerNorm"
        ln = (lambda k: g(k) + 1.0) if zero_centered else g          # TRT-LLM stores the plain gamma
        out["transformer.vocab_embedding.weight"] = g("embedding.word_embeddings.weight")
        if "embedding.position_embeddings.weight" in state_dict:
            out["transformer.position_embedding.weight"] = g("embedding.position_embeddings.weight")
        for i in range(c.num_layers):
            a, b = f"decoder.layers.{i}.", f"transformer.layers.{i}."
            for src, dst in (("input_layernorm", "input_layernorm"), ("pre_mlp_layernorm", "post_layernorm")):
                out[b + dst + ".weight"] = ln(a + src + ".weight")
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
            if a + "mlp.router.weight" in state_dict:
                self._convert_moe(state_dict, a, b, out, dtype)
                continue
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
        out["transformer.ln_f.weight"] = ln("decoder.final_layernorm.weight")
        if "decoder.final_layernorm.bias" in state_dict:
            out["transformer.ln_f.bias"] = g("decoder.final_layernorm.bias")
        out["lm_head.weight"] = g("output_layer.weight") if "output_layer.weight" in state_dict \
            else out["transformer.vocab_embedding.weight"]
        return out

    def _convert_moe(self, sd, a: str, b: str, out: Dict[str, torch.Tensor], dtype) -> None:
        """Experts are stacked along a leading axis: ``mlp.fc.weight`` [E, (2)ffn, h] with the *up* half first and the gate
        half second (TRT-LLM's MoE plugin order, the opposite of the dense gated MLP), ``mlp.proj.weight`` [E, h, ffn]; the
        router is kept in fp32."""
        c = self.cfg
        E = c.num_moe_experts
        out[b + "mlp.router.weight"] = sd[a + "mlp.router.weight"].float()
        if a + "mlp.experts.weight1" in sd:                                   # grouped layout: [E, h, (2)ffn] / [E, ffn, h]
            w1 = sd[a + "mlp.experts.weight1"].to(dtype).view(E, c.hidden_size, -1).transpose(1, 2)
            w2 = sd[a + "mlp.experts.weight2"].to(dtype).view(E, -1, c.hidden_size).transpose(1, 2)
        else:
            w1 = torch.stack([sd[a + f"mlp.experts.local_experts.{e}.linear_fc1.weight"].to(dtype) for e in range(E)])
            w2 = torch.stack([sd[a + f"mlp.experts.local_experts.{e}.linear_fc2.weight"].to(dtype) for e in range(E)])
        if c.gated_linear_unit:
            gate, up = torch.chunk(w1, 2, dim=1)
            w1 = torch.cat([up, gate], dim=1)
        out[b + "mlp.fc.weight"], out[b + "mlp.proj.weight"] = w1.contiguous(), w2.contiguous()

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
            elif w.dim() == 3 and k.endswith("mlp.fc.weight"):                # MoE: shard every expert's ffn axis
                if c.gated_linear_unit:
                    up, gate = torch.chunk(w, 2, dim=1)
                    out[k] = torch.cat([torch.chunk(up, tp, 1)[rank], torch.chunk(gate, tp, 1)[rank]], dim=1)
                else:
                    out[k] = torch.chunk(w, tp, dim=1)[rank]
            elif w.dim() == 3 and k.endswith("mlp.proj.weight"):
                out[k] = torch.chunk(w, tp, dim=2)[rank]
            elif k.endswith(("mlp.fc.weight", "mlp.gate.weight", "mlp.fc.bias", "lm_head.weight")):
                out[k] = torch.chunk(w, tp, dim=0)[rank]
            elif k.endswith(("attention.dense.weight", "mlp.proj.weight")):
                out[k] = torch.chunk(w, tp, dim=1)[rank]
            else:
                out[k] = w
        return {k: v.contiguous() for k, v in out.items()}

    # ---- pipeline split / vocabulary / fp8 -----------------------------------------------------------------------------
    def split_pipeline(self, weights: Dict[str, torch.Tensor], pp: int, pp_rank: int) -> Dict[str, torch.Tensor]:
        """Layers ``[pp_rank * L/pp, (pp_rank+1) * L/pp)`` keep their *global* index (TRT-LLM convention); the embedding
        lives on the first stage, the final norm and the head on the last."""
        if pp == 1:
            return weights
        L = self.cfg.num_layers
        if L % pp:
            raise ValueError(f"{L} layers do not split over {pp} pipeline stages")
        lo, hi = pp_rank * (L // pp), (pp_rank + 1) * (L // pp)
        out = {}
        for k, w in weights.items():
            if k.startswith("transformer.layers."):
                if lo <= int(k.split(".")[2]) < hi:
                    out[k] = w
            elif k.startswith(("transformer.vocab_embedding", "transformer.position_embedding")):
                if pp_rank == 0:
                    out[k] = w
            elif pp_rank == pp - 1:
                out[k] = w
        return out

    @staticmethod
    def remove_vocab_padding(weights: Dict[str, torch.Tensor], tokenizer_vocab_size: Optional[int]) -> Dict[str, torch.Tensor]:
        if tokenizer_vocab_size:
            for k in ("transformer.vocab_embedding.weight", "lm_head.weight"):
                if k in weights and weights[k].shape[0] > tokenizer_vocab_size:
                    weights[k] = weights[k][:tokenizer_vocab_size].contiguous()
        return weights

    def apply_fp8(self, weights: Dict[str, torch.Tensor], scales: Optional[Dict[str, Dict[str, torch.Tensor]]] = None,
                  fp8_kvcache: bool = False) -> Dict[str, torch.Tensor]:
        """GEMM weights of the transformer layers become ``float8_e4m3fn`` with a per-tensor ``weights_scaling_factor`` (amax /
        448 unless ``scales[<module>]`` carries calibrated ``{"weight_multiplier", "weights_scaling_factor",
        "activation_scaling_factor"}``), plus ``activation_scaling_factor`` and -- with ``fp8_kvcache`` -- unit
        ``kv_cache_scaling_factor`` entries, the keys TRT-LLM's FP8 quant mode loads."""
        scales = scales or {}
        out = dict(weights)
        gemms = ("attention.qkv", "attention.dense", "mlp.fc", "mlp.gate", "mlp.proj")
        for k, w in weights.items():
            mod = k.rsplit(".", 1)[0]
            if not (k.endswith(".weight") and k.startswith("transformer.layers.") and mod.endswith(gemms)) or w.dim() < 2:
                continue
            sc = scales.get(mod)
            if sc is not None:
                mult, wsf, asf = sc["weight_multiplier"], sc["weights_scaling_factor"], sc.get("activation_scaling_factor", torch.ones(1))
            else:
                amax = w.float().abs().max().clamp_min(1e-12)
                mult, wsf, asf = 448.0 / amax, (amax / 448.0).reshape(1), torch.ones(1)
            out[k] = (w.float() * mult).clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
            out[mod + ".weights_scaling_factor"] = torch.as_tensor(wsf, dtype=torch.float32).reshape(1)
            out[mod + ".activation_scaling_factor"] = torch.as_tensor(asf, dtype=torch.float32).reshape(1)
            if fp8_kvcache and mod.endswith("attention.qkv"):
                out[mod.rsplit(".", 1)[0] + ".kv_cache_scaling_factor"] = torch.ones(1, dtype=torch.float32)
        return out

    def get_trtllm_pretrained_config_and_model_weights(self, model_state_dict: Dict[str, torch.Tensor], dtype="bfloat16",
                                                       export_config=None, on_device_distributed_conversion: bool = False,
                                                       vocab_size: Optional[int] = None, fp8_quantized: bool = False,
                                                       fp8_kvcache: bool = False, scales=None):
        """Reference entry point (``trtllm_helper.py``): returns ``(weights per rank, config per rank)``.

        * default: ``model_state_dict`` is the full canonical model; one entry per ``tp x pp`` export rank;
        * ``on_device_distributed_conversion``: ``model_state_dict`` is THIS training rank's shard; one entry (this rank)."""
        from ..data_type import DataType
        from ..export_config import ExportConfig

        dt = DataType.of(dtype)
        if on_device_distributed_conversion:
            conv = DistributedTRTLLMConverter(self)
            w = conv.convert(model_state_dict, dt.torch_dtype, tokenizer_vocab_size=vocab_size)
            if fp8_quantized:
                w = self.apply_fp8(w, scales, fp8_kvcache)
            cfg = self.pretrained_config(vocab_size or conv.full_vocab, dt.value, conv.tp_size, conv.pp_size)
            cfg["mapping"]["rank"] = conv.pp_rank * conv.tp_size + conv.tp_rank
            self._quant_fields(cfg, fp8_quantized, fp8_kvcache)
            return [w], [cfg]
        ec = export_config or ExportConfig()
        full = self.remove_vocab_padding(self.convert(model_state_dict, dt.torch_dtype), vocab_size)
        if fp8_quantized:
            full = self.apply_fp8(full, scales, fp8_kvcache)
        weights, configs = [], []
        for pr in range(ec.inference_pp_size):
            stage = self.split_pipeline(full, ec.inference_pp_size, pr)
            for tr in range(ec.inference_tp_size):
                w = self.shard(stage, ec.inference_tp_size, tr)
                if ec.use_parallel_embedding and "transformer.vocab_embedding.weight" in w and ec.inference_tp_size > 1:
                    emb = w["transformer.vocab_embedding.weight"]
                    pad = (-emb.shape[0]) % ec.inference_tp_size
                    emb = torch.nn.functional.pad(emb, (0, 0, 0, pad)) if pad else emb
                    w["transformer.vocab_embedding.weight"] = torch.chunk(emb, ec.inference_tp_size, 0)[tr].contiguous()
                cfg = self.pretrained_config(full["transformer.vocab_embedding.weight"].shape[0], dt.value, ec.inference_tp_size,
                                             ec.inference_pp_size)
                cfg["use_parallel_embedding"] = bool(ec.use_parallel_embedding)
                cfg["mapping"]["rank"] = pr * ec.inference_tp_size + tr
                self._quant_fields(cfg, fp8_quantized, fp8_kvcache)
                weights.append(w)
                configs.append(cfg)
        return weights, configs

    @staticmethod
    def _quant_fields(cfg: Dict, fp8: bool, fp8_kv: bool) -> None:
        cfg["quantization"] = {"quant_algo": "FP8" if fp8 else None, "kv_cache_quant_algo": "FP8" if fp8_kv else None}

    def build_and_save_engine(self, engine_dir: str, trtllm_model_weights: Dict, trtllm_model_config: Dict, **kw):
        from ..trtllm_engine_builder import TRTLLMEngineBuilder

        return TRTLLMEngineBuilder.build_and_save_engine(engine_dir, trtllm_model_weights, trtllm_model_config, **kw)


class DistributedTRTLLMConverter:
    """On-device conversion of one training rank's shard (reference ``DistributedTRTLLMModelWeightsConverter``): the
    exported TP / PP layout *is* the training layout, so a column-parallel shard is already the right slice once its local
    QKV groups are regrouped to ``[q | k | v]`` and its local gated fc1 ``[gate_r ; up_r]`` is split; layer indices are
    shifted by the pipeline offset; the vocabulary-parallel embedding / head is summed over the TP group into the padded
    table, cut to the tokenizer size and -- for the head -- re-split."""

    def __init__(self, helper: TRTLLMHelper, tp_rank: Optional[int] = None, tp_size: Optional[int] = None, pp_rank: Optional[int] = None,
                 pp_size: Optional[int] = None, tp_all_reduce=None):
        from ...parallel import state as ps

        init = ps.model_parallel_is_initialized()
        self.h = helper
        self.tp_rank = tp_rank if tp_rank is not None else (ps.get_tensor_model_parallel_rank() if init else 0)
        self.tp_size = tp_size if tp_size is not None else (ps.get_tensor_model_parallel_world_size() if init else 1)
        self.pp_rank = pp_rank if pp_rank is not None else (ps.get_pipeline_model_parallel_rank() if init else 0)
        self.pp_size = pp_size if pp_size is not None else (ps.get_pipeline_model_parallel_world_size() if init else 1)
        if init and (ps.get_virtual_pipeline_model_parallel_world_size() or 1) > 1:
            raise ValueError("interleaved (virtual) pipeline chunks must be merged before an on-device export")
        self._all_reduce = tp_all_reduce or (lambda t: torch.distributed.all_reduce(t, group=ps.get_tensor_model_parallel_group()))
        self.full_vocab = None

    def _vocab(self, w: torch.Tensor, tokenizer_vocab_size: Optional[int], resplit: bool) -> torch.Tensor:
        tp, r = self.tp_size, self.tp_rank
        if tp > 1:
            full = torch.zeros(w.shape[0] * tp, *w.shape[1:], dtype=w.dtype, device=w.device)
            full[r * w.shape[0]:(r + 1) * w.shape[0]] = w
            self._all_reduce(full)
            w = full
        if tokenizer_vocab_size:
            w = w[:tokenizer_vocab_size]
        self.full_vocab = w.shape[0]
        if resplit and tp > 1:
            if w.shape[0] % tp:
                raise ValueError(f"tokenizer vocabulary {w.shape[0]} does not split over tp={tp} for the parallel head")
            w = torch.chunk(w, tp, 0)[r]
        return w.contiguous()

    @torch.no_grad()
    def convert(self, local_sd: Dict[str, torch.Tensor], dtype: torch.dtype = torch.bfloat16,
                tokenizer_vocab_size: Optional[int] = None) -> Dict[str, torch.Tensor]:
        import copy
        import re

        c = self.h.cfg
        tp = self.tp_size
        if c.num_layers % self.pp_size:
            raise ValueError("layers do not split evenly over the pipeline")
        per_stage = c.num_layers // self.pp_size
        local_cfg = copy.copy(c)                       # the single-device mapper sees a model with 1/tp of the heads / ffn
        local_cfg.num_attention_heads, local_cfg.num_query_groups = c.num_attention_heads // tp, max(1, c.num_query_groups // tp)
        local_cfg.num_layers = per_stage
        # local layer numbering -> 0..per_stage-1 for the mapper
        ids = sorted({int(m.group(1)) for k in local_sd for m in [re.match(r"decoder\.layers\.(\d+)\.", k)] if m})
        base = ids[0] if ids else 0
        sd = {}
        for k, v in local_sd.items():
            m = re.match(r"decoder\.layers\.(\d+)\.(.*)", k)
            sd[f"decoder.layers.{int(m.group(1)) - base}.{m.group(2)}" if m else k] = v
        first, last = self.pp_rank == 0, self.pp_rank == self.pp_size - 1
        placeholder = torch.zeros(1, c.hidden_size)
        sd.setdefault("embedding.word_embeddings.weight", placeholder)
        sd.setdefault("decoder.final_layernorm.weight", torch.zeros(c.hidden_size))
        local_helper = copy.copy(self.h)
        local_helper.cfg = local_cfg
        w = local_helper.convert(sd, dtype)
        out = {}
        off = self.pp_rank * per_stage
        for k, v in w.items():
            if k.startswith("transformer.layers."):
                parts = k.split(".")
                parts[2] = str(int(parts[2]) + off)
                out[".".join(parts)] = v
        if first:
            out["transformer.vocab_embedding.weight"] = self._vocab(local_sd["embedding.word_embeddings.weight"].to(dtype),
                                                                    tokenizer_vocab_size, resplit=False)
            if "embedding.position_embeddings.weight" in local_sd:
                out["transformer.position_embedding.weight"] = local_sd["embedding.position_embeddings.weight"].to(dtype)
        if last:
            for k in ("transformer.ln_f.weight", "transformer.ln_f.bias"):
                if k in w and "decoder.final_layernorm." + k.rsplit(".", 1)[1] in local_sd:
                    out[k] = w[k]
            head = local_sd.get("output_layer.weight")
            if head is None and "embedding.word_embeddings.weight" in local_sd:
                head = local_sd["embedding.word_embeddings.weight"]              # tied
            if head is not None:
                out["lm_head.weight"] = self._vocab(head.to(dtype), tokenizer_vocab_size, resplit=True)
        return out


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
