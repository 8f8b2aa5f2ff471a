"""TensorRT-LLM engine builder (reference ``core/export/trtllm/engine_builder/trtllm_engine_builder.py``).

The build *plan* -- plugin configuration, sequence / token limits (including TRT-LLM's ``max_num_tokens`` defaulting
rules), LoRA and prompt-table settings -- is computed here without TensorRT so it can be validated and stored next to the
exported checkpoint (``build_config.json``); ``build_and_save_engine`` then hands it to ``tensorrt_llm`` when that package
is importable (it is not part of this image: the call raises after writing the plan, and ``trtllm-build`` can consume the
checkpoint directory elsewhere)."""
from __future__ import annotations

import json
import math
import os
from dataclasses import asdict, dataclass, field
from typing import Dict, List, Optional


@dataclass
class EnginePlan:
    architecture: str
    max_input_len: int = 1024
    max_output_len: int = 1024
    max_batch_size: int = 4
    max_beam_width: int = 1
    max_seq_len: Optional[int] = None
    max_num_tokens: Optional[int] = None
    opt_num_tokens: Optional[int] = None
    max_prompt_embedding_table_size: int = 0
    use_refit: bool = False
    multiple_profiles: bool = False
    strongly_typed: bool = False
    gather_context_logits: bool = False
    gather_generation_logits: bool = False
    plugin: Dict = field(default_factory=dict)
    lora: Optional[Dict] = None

    def build_dict(self) -> Dict:
        d = asdict(self)
        d.pop("plugin")
        d.pop("lora")
        d.pop("architecture")
        d["builder_opt"] = None
        return d


def resolve_num_tokens(max_num_tokens, opt_num_tokens, max_seq_len, max_batch_size, max_input_len, max_beam_width,
                       remove_input_padding, enable_context_fmha, tokens_per_block, multiple_profiles):
    """TRT-LLM's defaulting of the packed-token limits: without input-padding removal the limits are unused; otherwise
    ``max_num_tokens`` defaults to ``max_batch_size * max_seq_len`` (capped to ``max_batch_size * max_input_len`` worth of
    context), ``opt_num_tokens`` to ``max_batch_size * max_beam_width`` (pure generation); with paged context attention the
    limit is rounded up to a whole KV block."""
    if not remove_input_padding:
        return None, None
    if max_num_tokens is None:
        max_num_tokens = max_seq_len * max_batch_size
    max_num_tokens = min(max_num_tokens, max_seq_len * max_batch_size)
    if max_num_tokens < max_input_len and not enable_context_fmha:
        max_num_tokens = max_input_len                   # a whole prompt must fit in one context step
    if enable_context_fmha and max_num_tokens % tokens_per_block:
        max_num_tokens = int(math.ceil(max_num_tokens / tokens_per_block)) * tokens_per_block
    if opt_num_tokens is None and not multiple_profiles:
        opt_num_tokens = min(max_batch_size * max_beam_width, max_num_tokens)
    if opt_num_tokens is not None and opt_num_tokens > max_num_tokens:
        opt_num_tokens = max_num_tokens
    return max_num_tokens, opt_num_tokens


class TRTLLMEngineBuilder:
    @staticmethod
    def plan(trtllm_model_config: Dict, max_input_len: int = 1024, max_output_len: int = 1024, max_batch_size: int = 4,
             lora_ckpt_list: Optional[List[str]] = None, use_lora_plugin: Optional[str] = None, max_lora_rank: int = 64,
             lora_target_modules: Optional[List[str]] = None, max_prompt_embedding_table_size: int = 0,
             paged_kv_cache: bool = True, remove_input_padding: bool = True, paged_context_fmha: bool = False,
             use_refit: bool = False, max_num_tokens: Optional[int] = None, max_seq_len: Optional[int] = None,
             opt_num_tokens: Optional[int] = None, max_beam_width: int = 1, tokens_per_block: int = 128,
             multiple_profiles: bool = False, gpt_attention_plugin: str = "auto", gemm_plugin: str = "auto",
             reduce_fusion: bool = False) -> EnginePlan:
        arch = trtllm_model_config["architecture"]
        if max_seq_len is None:
            max_seq_len = max_input_len + max_output_len
        if max_input_len > max_seq_len:
            raise ValueError(f"max_input_len {max_input_len} exceeds max_seq_len {max_seq_len}")
        limit = trtllm_model_config.get("max_position_embeddings")
        if limit and trtllm_model_config.get("position_embedding_type") == "learned_absolute" and max_seq_len > limit:
            raise ValueError(f"max_seq_len {max_seq_len} exceeds the learned position table ({limit})")
        plugin = {"gpt_attention_plugin": gpt_attention_plugin, "gemm_plugin": gemm_plugin, "paged_kv_cache": bool(paged_kv_cache),
                  "tokens_per_block": tokens_per_block if paged_kv_cache else None, "remove_input_padding": remove_input_padding,
                  "use_paged_context_fmha": paged_context_fmha, "multiple_profiles": multiple_profiles,
                  "reduce_fusion": reduce_fusion, "context_fmha": True, "use_fused_mlp": True}
        if paged_context_fmha and not paged_kv_cache:
            raise ValueError("paged_context_fmha needs paged_kv_cache")
        mnt, ont = resolve_num_tokens(max_num_tokens, opt_num_tokens, max_seq_len, max_batch_size, max_input_len, max_beam_width,
                                      remove_input_padding, plugin["context_fmha"] and paged_context_fmha, tokens_per_block,
                                      multiple_profiles)
        plan = EnginePlan(architecture=arch, max_input_len=max_input_len, max_output_len=max_output_len, max_batch_size=max_batch_size,
                          max_beam_width=max_beam_width, max_seq_len=max_seq_len, max_num_tokens=mnt, opt_num_tokens=ont,
                          max_prompt_embedding_table_size=max_prompt_embedding_table_size, use_refit=use_refit,
                          multiple_profiles=multiple_profiles, plugin=plugin)
        if arch == "DeciLMForCausalLM":                     # heterogeneous (NAS) models: per-layer shapes need strong typing
            plan.strongly_typed = True
            plan.plugin["use_fused_mlp"] = False
        if use_lora_plugin is not None:
            plan.plugin["lora_plugin"] = use_lora_plugin
            plan.lora = {"lora_dir": lora_ckpt_list, "lora_ckpt_source": "nemo", "max_lora_rank": max_lora_rank,
                         "lora_target_modules": lora_target_modules}
        return plan

    @staticmethod
    def build_and_save_engine(engine_dir: str, trtllm_model_weights: Dict, trtllm_model_config: Dict, **kw):
        plan = TRTLLMEngineBuilder.plan(trtllm_model_config, **kw)
        os.makedirs(engine_dir, exist_ok=True)
        with open(os.path.join(engine_dir, "build_config.json"), "w") as f:
            json.dump({"architecture": plan.architecture, "build": plan.build_dict(), "plugin_config": plan.plugin,
                       "lora_config": plan.lora}, f, indent=1)
        try:
            import tensorrt_llm
        except ImportError as e:
            raise RuntimeError(f"tensorrt_llm is not installed; the build plan was written to {engine_dir}/build_config.json -- "
                               "run `trtllm-build --checkpoint_dir <exported checkpoint>` with these limits where it is") from e
        from tensorrt_llm.builder import BuildConfig
        from tensorrt_llm.commands.build import build as build_trtllm
        from tensorrt_llm.models.modeling_utils import PretrainedConfig, optimize_model, preprocess_weights
        from tensorrt_llm.plugin import PluginConfig

        arch = "LLaMAForCausalLM" if plan.architecture == "LlamaForCausalLM" else plan.architecture
        model_cls = getattr(tensorrt_llm.models, arch)
        pc = PluginConfig()
        for k, v in plan.plugin.items():
            if k == "tokens_per_block":
                continue
            if hasattr(pc, k):
                setattr(pc, k, v)
        if plan.plugin["paged_kv_cache"]:
            pc.enable_paged_kv_cache(tokens_per_block=plan.plugin["tokens_per_block"])
        build_config = BuildConfig.from_dict(plan.build_dict(), plugin_config=pc)
        if plan.lora is not None:
            from tensorrt_llm.lora_manager import LoraConfig

            build_config.lora_config = LoraConfig(**plan.lora)
        cfg = PretrainedConfig.from_dict(trtllm_model_config)
        model = optimize_model(model_cls.from_config(cfg), use_parallel_embedding=cfg.use_parallel_embedding,
                               share_embedding_table=cfg.share_embedding_table)
        preprocess_weights(trtllm_model_weights, cfg)
        model.load(trtllm_model_weights)
        engine = build_trtllm(model, build_config)
        engine.save(engine_dir)
        return engine
