"""Drop-in ``megatron`` namespace: the import paths of Megatron-LM / MegatronApp resolved onto the
B200-native implementation, so existing user scripts (``from megatron.core import mpu,
tensor_parallel``, ``from megatron.training import get_args, pretrain`` ...) keep working.

It lives under ``compat/`` (add that directory to ``PYTHONPATH``) rather than at the repository root so
that it never shadows a real Megatron checkout -- in particular the unmodified reference that the
benchmark's ``--impl reference`` arm imports from ``baseline/_ref``.

Only aliases live here -- no logic.  ``import megatron`` installs lazy module aliases in
``sys.modules``; the real code is in ``megatronapp_b200``.
"""
import importlib
import sys
import types

_ALIASES = {
    # megatron.core
    "megatron.core.parallel_state": "megatronapp_b200.parallel.state",
    "megatron.core.mpu": "megatronapp_b200.parallel.state",
    "megatron.core.model_parallel_config": "megatronapp_b200.config.transformer_config",
    "megatron.core.transformer.transformer_config": "megatronapp_b200.config.transformer_config",
    "megatron.core.transformer.spec_utils": "megatronapp_b200.transformer.spec_utils",
    "megatron.core.transformer.module": "megatronapp_b200.transformer.module",
    "megatron.core.transformer.attention": "megatronapp_b200.transformer.attention",
    "megatron.core.transformer.dot_product_attention": "megatronapp_b200.transformer.attention",
    "megatron.core.transformer.multi_latent_attention": "megatronapp_b200.transformer.mla",
    "megatron.core.transformer.multi_token_prediction": "megatronapp_b200.transformer.multi_token_prediction",
    "megatron.core.transformer.mlp": "megatronapp_b200.transformer.mlp",
    "megatron.core.transformer.transformer_layer": "megatronapp_b200.transformer.transformer_layer",
    "megatron.core.transformer.transformer_block": "megatronapp_b200.transformer.transformer_block",
    "megatron.core.transformer.cuda_graphs": "megatronapp_b200.transformer.cuda_graphs",
    "megatron.core.transformer.torch_norm": "megatronapp_b200.transformer.norms",
    "megatron.core.transformer.moe.moe_layer": "megatronapp_b200.moe.moe_layer",
    "megatron.core.transformer.moe.router": "megatronapp_b200.moe.router",
    "megatron.core.transformer.moe.experts": "megatronapp_b200.moe.experts",
    "megatron.core.transformer.moe.moe_utils": "megatronapp_b200.moe.moe_utils",
    "megatron.core.transformer.moe.token_dispatcher": "megatronapp_b200.moe.token_dispatcher",
    "megatron.core.transformer.moe.shared_experts": "megatronapp_b200.moe.shared_experts",
    "megatron.core.transformer.moe.upcycling_utils": "megatronapp_b200.moe.upcycling_utils",
    "megatron.core.fusions.fused_layer_norm": "megatronapp_b200.transformer.norms",
    "megatron.core.tensor_parallel.layers": "megatronapp_b200.parallel.tp_layers",
    "megatron.core.tensor_parallel.mappings": "megatronapp_b200.parallel.mappings",
    "megatron.core.tensor_parallel.cross_entropy": "megatronapp_b200.parallel.cross_entropy",
    "megatron.core.tensor_parallel.random": "megatronapp_b200.parallel.random",
    "megatron.core.pipeline_parallel.schedules": "megatronapp_b200.parallel.schedules",
    "megatron.core.pipeline_parallel.p2p_communication": "megatronapp_b200.parallel.p2p",
    "megatron.core.distributed.distributed_data_parallel": "megatronapp_b200.parallel.ddp",
    "megatron.core.distributed.custom_fsdp": "megatronapp_b200.parallel.fsdp",
    "megatron.core.distributed.torch_fully_sharded_data_parallel": "megatronapp_b200.parallel.fsdp",
    "megatron.core.optimizer.optimizer": "megatronapp_b200.optim.optimizer",
    "megatron.core.optimizer_param_scheduler": "megatronapp_b200.optim.scheduler",
    "megatron.core.num_microbatches_calculator": "megatronapp_b200.utils.microbatches",
    "megatron.core.timers": "megatronapp_b200.utils.timers",
    "megatron.core.rerun_state_machine": "megatronapp_b200.training.rerun_state_machine",
    "megatron.core.dist_checkpointing.serialization": "megatronapp_b200.ckpt.serialization",
    "megatron.core.dist_checkpointing.validation": "megatronapp_b200.ckpt.serialization",
    "megatron.core.dist_checkpointing.core": "megatronapp_b200.ckpt.serialization",
    "megatron.core.dist_checkpointing.mapping": "megatronapp_b200.ckpt",
    "megatron.core.models.gpt.gpt_model": "megatronapp_b200.models.gpt.gpt_model",
    "megatron.core.models.gpt.gpt_layer_specs": "megatronapp_b200.models.gpt.gpt_layer_specs",
    "megatron.core.models.bert.bert_model": "megatronapp_b200.models.bert.bert_model",
    "megatron.core.models.T5.t5_model": "megatronapp_b200.models.T5.t5_model",
    "megatron.core.models.mamba.mamba_model": "megatronapp_b200.models.mamba.mamba_model",
    "megatron.core.models.common.embeddings": "megatronapp_b200.models.common.embeddings",
    "megatron.core.ssm.mamba_mixer": "megatronapp_b200.ssm.mamba",
    "megatron.core.datasets.indexed_dataset": "megatronapp_b200.data.indexed_dataset",
    "megatron.core.datasets.gpt_dataset": "megatronapp_b200.data.gpt_dataset",
    "megatron.core.datasets.helpers": "megatronapp_b200.data._helpers",
    "megatron.core.datasets.blended_megatron_dataset_builder": "megatronapp_b200.data.blended_builder",
    "megatron.core.datasets.blended_megatron_dataset_config": "megatronapp_b200.data.blended_builder",
    "megatron.core.datasets.masked_dataset": "megatronapp_b200.data.masked_dataset",
    "megatron.core.datasets.bert_dataset": "megatronapp_b200.data.masked_dataset",
    "megatron.core.datasets.t5_dataset": "megatronapp_b200.data.masked_dataset",
    "megatron.core.packed_seq_params": "megatronapp_b200.utils.packed_seq_params",
    "megatron.core.process_groups_config": "megatronapp_b200.parallel.process_groups_config",
    "megatron.core.fp8_utils": "megatronapp_b200.ops.fp8",
    "megatron.core.inference.engines": "megatronapp_b200.inference",
    "megatron.core.inference.sampling_params": "megatronapp_b200.inference.core",
    "megatron.core.inference.inference_request": "megatronapp_b200.inference.core",
    "megatron.core.inference.scheduler": "megatronapp_b200.inference.core",
    "megatron.core.inference.async_stream": "megatronapp_b200.inference.async_stream",
    "megatron.core.inference.common_inference_params": "megatronapp_b200.inference.core",
    "megatron.core.inference.engines.dynamic_engine": "megatronapp_b200.inference.dynamic",
    "megatron.core.inference.engines.static_engine": "megatronapp_b200.inference.core",
    "megatron.core.inference.contexts.dynamic_context": "megatronapp_b200.inference.dynamic",
    "megatron.core.inference.contexts.static_context": "megatronapp_b200.inference.core",
    "megatron.core.inference.contexts": "megatronapp_b200.inference",
    "megatron.core.inference.text_generation_controllers.text_generation_controller": "megatronapp_b200.inference.core",
    "megatron.core.models.vision.clip_vit_model": "megatronapp_b200.models.vision.clip_vit_model",
    "megatron.core.models.vision.radio": "megatronapp_b200.models.vision.radio",
    "megatron.core.models.vision.multimodal_projector": "megatronapp_b200.models.vision.multimodal_projector",
    "megatron.core.models.multimodal.llava_model": "megatronapp_b200.models.multimodal.llava_model",
    "megatron.core.models.multimodal.llava_spec": "megatronapp_b200.models.multimodal.llava_spec",
    "megatron.core.models.multimodal.context_parallel": "megatronapp_b200.models.multimodal.context_parallel",
    "megatron.core.models.vision.vit_layer_specs": "megatronapp_b200.models.vision.vit_layer_specs",
    "megatron.core.models.huggingface": "megatronapp_b200.models.huggingface",
    "megatron.core.models.huggingface.module": "megatronapp_b200.models.huggingface.module",
    "megatron.core.models.huggingface.clip_model": "megatronapp_b200.models.huggingface.clip_model",
    "megatron.core.models.huggingface.qwen_model": "megatronapp_b200.models.huggingface.qwen_model",
    "megatron.core.models.common.vision_module.vision_module": "megatronapp_b200.models.common.vision_module.vision_module",
    "megatron.core.models.retro": "megatronapp_b200.models.retro",
    "megatron.core.models.common.embeddings.relative_pos_embedding": "megatronapp_b200.models.common.embeddings",
    "megatron.core.models.common.embeddings.rotary_pos_embedding": "megatronapp_b200.models.common.embeddings",
    "megatron.core.models.common.language_module.language_module": "megatronapp_b200.models.common.language_module",
    "megatron.core.export.trtllm": "megatronapp_b200.export.trtllm",
    "megatron.core.export.trtllm.trtllm_helper": "megatronapp_b200.export.trtllm.trtllm_helper",
    "megatron.core.export.trtllm.engine_builder": "megatronapp_b200.export.trtllm.engine_builder",
    "megatron.core.export.trtllm.engine_builder.trtllm_engine_builder": "megatronapp_b200.export.trtllm.engine_builder.trtllm_engine_builder",
    "megatron.core.export.trtllm.trtllm_weights_converter": "megatronapp_b200.export.trtllm.trtllm_weights_converter",
    "megatron.core.export.data_type": "megatronapp_b200.export.data_type",
    "megatron.core.export.export_config": "megatronapp_b200.export.export_config",
    "megatron.core.export.model_type": "megatronapp_b200.export.model_type",
    "megatron.core.datasets.retro": "megatronapp_b200.data.retro",
    "megatron.core.datasets.retro.config": "megatronapp_b200.data.retro.config",
    "megatron.core.datasets.retro.db": "megatronapp_b200.data.retro.db",
    "megatron.core.datasets.retro.index": "megatronapp_b200.data.retro.index",
    "megatron.core.datasets.retro.query": "megatronapp_b200.data.retro.query",
    "megatron.legacy.model.vision": "megatronapp_b200.legacy.model.vision",
    "megatron.training.local_checkpoint": "megatronapp_b200.training.local_checkpoint",
    "megatron.core.post_training.modelopt": "megatronapp_b200.post_training.quantization",
    "megatron.legacy.model": "megatronapp_b200.legacy.model",
    "megatron.legacy.model.biencoder_model": "megatronapp_b200.legacy.model.biencoder_model",
    "megatron.legacy.data.ict_dataset": "megatronapp_b200.data.ict_dataset",
    "megatron.training.tokenizer.multimodal_tokenizer": "megatronapp_b200.data.tokenizer_extra",
    "megatron.training.yaml_arguments": "megatronapp_b200.training.yaml_arguments",
    "megatron.training.ft_integration": "megatronapp_b200.training.ft_integration",
    "megatron.training.async_utils": "megatronapp_b200.training.async_utils",
    "megatron.post_training.checkpointing": "megatronapp_b200.post_training.quantization",
    "megatron.post_training.algos.distillation": "megatronapp_b200.post_training.distillation",
    # megatron.training & add-ons
    "megatron.training.arguments": "megatronapp_b200.training.arguments",
    "megatron.training.initialize": "megatronapp_b200.training.initialize",
    "megatron.training.global_vars": "megatronapp_b200.training.global_vars",
    "megatron.training.training": "megatronapp_b200.training.training",
    "megatron.training.checkpointing": "megatronapp_b200.training.checkpointing",
    "megatron.training.utils": "megatronapp_b200.training.utils",
    "megatron.training.trace": "megatronapp_b200.scan.tracer",
    "megatron.training.theoretical_memory_usage": "megatronapp_b200.training.theoretical_memory_usage",
    "megatron.training.tokenizer.tokenizer": "megatronapp_b200.data.tokenizer",
    "megatron.training.training_wsserver": "megatronapp_b200.scope.training_server",
    "megatron.core.tensor_tracer": "megatronapp_b200.scope.capture",
    "megatron.core.tensor_disturbance": "megatronapp_b200.scope.disturbance",
    "megatron.inference.text_generation_server": "megatronapp_b200.inference.server",
    "megatron.inference.text_generation": "megatronapp_b200.inference.text_generation",
    "megatron.inference.text_generation.api": "megatronapp_b200.inference.text_generation.api",
    "megatron.inference.text_generation.beam_utils": "megatronapp_b200.inference.beam_search",
    "megatron.legacy.data.data_samplers": "megatronapp_b200.data.samplers",
}

_PACKAGES = {
    "megatron.core": ("megatronapp_b200", {"parallel_state": "megatronapp_b200.parallel.state",
                                           "mpu": "megatronapp_b200.parallel.state",
                                           "tensor_parallel": None, "ModelParallelConfig": None}),
}


class _Lazy(types.ModuleType):
    """Namespace module whose attributes resolve to aliased modules on first access."""

    def __init__(self, name, doc=None):
        super().__init__(name, doc)
        self.__path__ = []          # mark as package so "import a.b.c" walks through it

    def __getattr__(self, item):
        full = f"{self.__name__}.{item}"
        target = _ALIASES.get(full)
        if target is not None:
            mod = importlib.import_module(target)
            sys.modules[full] = mod
            setattr(self, item, mod)
            return mod
        if any(k.startswith(full + ".") for k in _ALIASES):
            mod = _ensure_pkg(full)
            setattr(self, item, mod)
            return mod
        for src in _UNIONS.get(self.__name__, ()):
            m = importlib.import_module(src)
            if hasattr(m, item):
                val = getattr(m, item)
                setattr(self, item, val)
                return val
        extra = _EXTRA_ATTRS.get(self.__name__, {}).get(item)
        if extra is not None:
            val = extra()
            setattr(self, item, val)
            return val
        raise AttributeError(f"module {self.__name__!r} has no attribute {item!r}")


def _ensure_pkg(name):
    mod = sys.modules.get(name)
    if mod is None:
        mod = _Lazy(name)
        sys.modules[name] = mod
    return mod


def _attr(module, name):
    return lambda: getattr(importlib.import_module(module), name)


# flat namespaces that re-export several native modules (``tensor_parallel.ColumnParallelLinear`` ...)
_UNIONS = {
    "megatron.core.tensor_parallel": ("megatronapp_b200.parallel.tp_layers", "megatronapp_b200.parallel.mappings",
                                      "megatronapp_b200.parallel.cross_entropy", "megatronapp_b200.parallel.random"),
    "megatron.core.pipeline_parallel": ("megatronapp_b200.parallel.schedules",),
    "megatron.core.datasets": ("megatronapp_b200.data.gpt_dataset", "megatronapp_b200.data.indexed_dataset",
                               "megatronapp_b200.data.blended_builder", "megatronapp_b200.data.mock"),
    "megatron.core.datasets.gpt_dataset": (),
    "megatron.core.dist_checkpointing": ("megatronapp_b200.ckpt", "megatronapp_b200.ckpt.sharded"),
    "megatron.core.inference": ("megatronapp_b200.inference", "megatronapp_b200.inference.core"),
    "megatron.training.tokenizer": ("megatronapp_b200.data.tokenizer", "megatronapp_b200.data.tokenizer_extra"),
}


_EXTRA_ATTRS = {
    "megatron.core": {
        "ModelParallelConfig": _attr("megatronapp_b200.config.transformer_config", "ModelParallelConfig"),
    },
    "megatron.training": {k: _attr(m, k) for k, m in {
        "get_args": "megatronapp_b200.training.global_vars", "get_timers": "megatronapp_b200.training.global_vars",
        "get_tokenizer": "megatronapp_b200.training.global_vars", "get_tracer": "megatronapp_b200.scan.tracer",
        "get_tensorboard_writer": "megatronapp_b200.training.global_vars",
        "pretrain": "megatronapp_b200.training.training", "print_rank_0": "megatronapp_b200.training.utils",
        "initialize_megatron": "megatronapp_b200.training.initialize"}.items()},
    "megatron.core.transformer": {
        "TransformerConfig": _attr("megatronapp_b200.config.transformer_config", "TransformerConfig"),
        "MLATransformerConfig": _attr("megatronapp_b200.config.transformer_config", "MLATransformerConfig"),
        "ModuleSpec": _attr("megatronapp_b200.transformer.spec_utils", "ModuleSpec"),
        "MegatronModule": _attr("megatronapp_b200.transformer.module", "MegatronModule"),
    },
    "megatron.core.models.gpt": {"GPTModel": _attr("megatronapp_b200.models.gpt.gpt_model", "GPTModel")},
    "megatron.core.distributed": {
        "DistributedDataParallel": _attr("megatronapp_b200.parallel.ddp", "DistributedDataParallel"),
        "DistributedDataParallelConfig": _attr("megatronapp_b200.parallel.ddp", "DistributedDataParallelConfig"),
        "finalize_model_grads": _attr("megatronapp_b200.parallel.ddp", "finalize_model_grads"),
    },
    "megatron.core.optimizer": {
        "OptimizerConfig": _attr("megatronapp_b200.optim", "OptimizerConfig"),
        "get_megatron_optimizer": _attr("megatronapp_b200.optim", "get_megatron_optimizer"),
    },
    "megatron.core.pipeline_parallel": {
        "get_forward_backward_func": _attr("megatronapp_b200.parallel.schedules", "get_forward_backward_func"),
    },
}


class _Finder:
    """Resolves ``import megatron.x.y`` for aliased leaves and synthesises intermediate packages."""

    @staticmethod
    def find_spec(name, path=None, target=None):
        if not name.startswith("megatron."):
            return None
        from importlib.machinery import ModuleSpec

        if name in _ALIASES or any(k.startswith(name + ".") for k in _ALIASES) or name in _EXTRA_ATTRS:
            return ModuleSpec(name, _Finder, is_package=name not in _ALIASES)
        return None

    @staticmethod
    def create_module(spec):
        if spec.name in _ALIASES:
            return importlib.import_module(_ALIASES[spec.name])
        return _Lazy(spec.name)

    @staticmethod
    def exec_module(module):
        return None


if not any(isinstance(f, type) and f.__name__ == "_Finder" for f in sys.meta_path):
    sys.meta_path.insert(0, _Finder)
