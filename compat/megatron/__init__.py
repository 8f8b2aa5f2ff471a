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
    # function-level API (megatronapp_b200.core_api) and further module paths of the reference
    "megatron.core.enums": "megatronapp_b200.core_api.enums",
    "megatron.core.transformer.enums": "megatronapp_b200.core_api.enums",
    "megatron.legacy.model.enums": "megatronapp_b200.core_api.enums",
    "megatron.core.optimizer.clip_grads": "megatronapp_b200.core_api.clip_grads",
    "megatron.core.tensor_parallel.utils": "megatronapp_b200.core_api.tp_utils",
    "megatron.core.tensor_parallel.data": "megatronapp_b200.parallel.data",
    "megatron.core.fusions.fused_bias_gelu": "megatronapp_b200.core_api.fusions",
    "megatron.core.fusions.fused_bias_swiglu": "megatronapp_b200.core_api.fusions",
    "megatron.core.fusions.fused_bias_geglu": "megatronapp_b200.core_api.fusions",
    "megatron.core.fusions.fused_bias_dropout": "megatronapp_b200.core_api.fusions",
    "megatron.core.fusions.fused_softmax": "megatronapp_b200.core_api.fusions",
    "megatron.core.fusions.fused_cross_entropy": "megatronapp_b200.core_api.fusions",
    "megatron.legacy.model.fused_bias_gelu": "megatronapp_b200.core_api.fusions",
    "megatron.legacy.model.fused_softmax": "megatronapp_b200.core_api.fusions",
    "megatron.core.extensions.transformer_engine": "megatronapp_b200.core_api.te",
    "megatron.core.transformer.custom_layers.transformer_engine": "megatronapp_b200.core_api.te",
    "megatron.core.inference_params": "megatronapp_b200.inference.core",
    "megatron.core.inference.model_inference_wrappers.gpt.gpt_inference_wrapper": "megatronapp_b200.inference.core",
    "megatron.core.inference.model_inference_wrappers.t5.t5_inference_wrapper": "megatronapp_b200.inference.core",
    "megatron.core.inference.model_inference_wrappers.multimodal.vlm_inference_wrapper": "megatronapp_b200.inference.core",
    "megatron.core.inference.model_inference_wrappers.abstract_model_inference_wrapper": "megatronapp_b200.inference.core",
    "megatron.core.inference.engines.mcore_engine": "megatronapp_b200.inference.core",
    "megatron.core.inference.engines.abstract_engine": "megatronapp_b200.inference.core",
    "megatron.core.inference.text_generation_controllers.simple_text_generation_controller": "megatronapp_b200.inference.core",
    "megatron.core.inference.text_generation_controllers.encoder_decoder_text_generation_controller": "megatronapp_b200.inference.core",
    "megatron.core.inference.text_generation_controllers.vlm_text_generation_controller": "megatronapp_b200.inference.core",
    "megatron.core.inference.communication_utils": "megatronapp_b200.core_api.text_generation",
    "megatron.inference.text_generation.sampling": "megatronapp_b200.core_api.text_generation",
    "megatron.inference.text_generation.tokenization": "megatronapp_b200.core_api.text_generation",
    "megatron.inference.text_generation.communication": "megatronapp_b200.core_api.text_generation",
    "megatron.inference.text_generation.generation": "megatronapp_b200.inference.text_generation.api",
    "megatron.inference.endpoints.completions": "megatronapp_b200.inference.server",
    "megatron.inference.endpoints.common": "megatronapp_b200.inference.server",
    "megatron.core.optimizer.optimizer_config": "megatronapp_b200.optim.optimizer",
    "megatron.core.optimizer.distrib_optimizer": "megatronapp_b200.optim.optimizer",
    "megatron.core.optimizer.grad_scaler": "megatronapp_b200.optim.optimizer",
    "megatron.core.distributed.distributed_data_parallel_config": "megatronapp_b200.parallel.ddp",
    "megatron.core.distributed.finalize_model_grads": "megatronapp_b200.parallel.ddp",
    "megatron.core.distributed.param_and_grad_buffer": "megatronapp_b200.parallel.ddp",
    "megatron.core.distributed.data_parallel_base": "megatronapp_b200.parallel.ddp",
    "megatron.core.distributed.custom_fsdp.fully_sharded_data_parallel": "megatronapp_b200.parallel.fsdp",
    "megatron.core.models.retro.config": "megatronapp_b200.models.retro.retro_model",
    "megatron.core.models.retro.model": "megatronapp_b200.models.retro.retro_model",
    "megatron.core.models.retro.decoder_spec": "megatronapp_b200.models.retro.retro_model",
    "megatron.core.models.retro.encoder_spec": "megatronapp_b200.models.retro.retro_model",
    "megatron.core.models.bert.bert_layer_specs": "megatronapp_b200.models.bert.bert_model",
    "megatron.core.models.bert.bert_lm_head": "megatronapp_b200.models.bert.bert_model",
    "megatron.core.models.bert.pooler": "megatronapp_b200.models.bert.bert_model",
    "megatron.core.models.common.embeddings.language_model_embedding": "megatronapp_b200.models.common.embeddings",
    "megatron.core.models.common.embeddings.rotary_pos_embedding": "megatronapp_b200.models.common.embeddings",
    "megatron.core.models.common.embeddings.relative_pos_embedding": "megatronapp_b200.models.common.embeddings",
    "megatron.core.models.common.embeddings.yarn_rotary_pos_embedding": "megatronapp_b200.models.common.embeddings",
    "megatron.core.models.gpt.heterogeneous.heterogeneous_layer_specs": "megatronapp_b200.models.gpt.heterogeneous.heterogeneous_layer_specs",
    "megatron.core.ssm.mamba_block": "megatronapp_b200.ssm.mamba",
    "megatron.core.ssm.mamba_layer": "megatronapp_b200.ssm.mamba",
    "megatron.core.ssm.mamba_hybrid_layer_allocation": "megatronapp_b200.ssm.mamba",
    "megatron.core.transformer.identity_op": "megatronapp_b200.transformer.norms",
    "megatron.core.transformer.torch_layer_norm": "megatronapp_b200.transformer.norms",
    "megatron.core.config_logger": "megatronapp_b200.utils.config_logger",
    "megatron.core.trace": "megatronapp_b200.scan.tracer",
    "megatron.core.datasets.megatron_tokenizer": "megatronapp_b200.data.tokenizer",
    "megatron.core.datasets.blended_dataset": "megatronapp_b200.data.gpt_dataset",
    "megatron.core.datasets.retro.config.config": "megatronapp_b200.data.retro.config",
    "megatron.core.datasets.retro.query.retro_dataset": "megatronapp_b200.data.retro.query",
    "megatron.core.datasets.retro.query.query": "megatronapp_b200.data.retro.query",
    "megatron.core.datasets.retro.db.build": "megatronapp_b200.data.retro.db",
    "megatron.core.datasets.retro.index.index": "megatronapp_b200.data.retro.index",
    "megatron.core.export.trtllm.trt_model_type": "megatronapp_b200.export.model_type",
    "megatron.legacy.data.autoaugment": "megatronapp_b200.legacy.data.autoaugment",
    "megatron.legacy.data.image_folder": "megatronapp_b200.legacy.data.image_folder",
    "megatron.legacy.data.vit_dataset": "megatronapp_b200.legacy.data.vit_dataset",
    "megatron.legacy.data.realm_index": "megatronapp_b200.legacy.data.realm_index",
    "megatron.legacy.data.orqa_wiki_dataset": "megatronapp_b200.legacy.data.orqa_wiki_dataset",
    "megatron.legacy.data.multimodal_dataset": "megatronapp_b200.legacy.data.multimodal_dataset",
    "megatron.legacy.indexer": "megatronapp_b200.legacy.indexer",
    "megatron.legacy.model.bert_model": "megatronapp_b200.models.bert.bert_model",
    "megatron.legacy.model.classification": "megatronapp_b200.models.bert.classification",
    "megatron.legacy.model.multiple_choice": "megatronapp_b200.models.bert.classification",
    "megatron.legacy.model.module": "megatronapp_b200.transformer.module",
    "megatron.legacy.model.vision.esvit_swin_backbone": "megatronapp_b200.legacy.model.vision.swin_backbone",
    "megatron.legacy.model.vision.classification": "megatronapp_b200.models.vision.vit_classification",
    "megatron.training.dist_signal_handler": "megatronapp_b200.training.signals",
    "megatron.training.one_logger_utils": "megatronapp_b200.training.one_logger_utils",
    "megatron.training.wandb_utils": "megatronapp_b200.training.wandb_utils",
    "megatron.training.activations": "megatronapp_b200.transformer.mlp",
    "megatron.post_training.arguments": "megatronapp_b200.training.arguments",
    "megatron.core.models.common.embeddings.rope_utils": "megatronapp_b200.models.common.embeddings",
    "megatron.core.models.mamba.mamba_layer_specs": "megatronapp_b200.models.mamba.mamba_model",
    "megatron.core.models.T5.t5_spec": "megatronapp_b200.models.T5.t5_model",
    "megatron.core.models.gpt.moe_module_specs": "megatronapp_b200.models.gpt.gpt_layer_specs",
    "megatron.core.models.retro.base_attention": "megatronapp_b200.models.retro.retro_model",
    "megatron.core.models.retro.decoder_attention": "megatronapp_b200.models.retro.retro_model",
    "megatron.core.models.retro.encoder_attention": "megatronapp_b200.models.retro.retro_model",
    "megatron.core.models.retro.utils": "megatronapp_b200.models.retro.retro_model",
    "megatron.core.transformer.heterogeneous.heterogeneous_config": "megatronapp_b200.models.gpt.heterogeneous.heterogeneous_layer_specs",
    "megatron.core.transformer.moe.legacy_a2a_token_dispatcher": "megatronapp_b200.moe.token_dispatcher",
    "megatron.core.transformer.moe.fused_a2a": "megatronapp_b200.moe.token_dispatcher",
    "megatron.core.transformer.moe.grouped_gemm_util": "megatronapp_b200.ops.moe",
    "megatron.core.distributed.custom_fsdp.param_and_grad_buffer": "megatronapp_b200.parallel.fsdp",
    "megatron.core.ssm.mamba_config": "megatronapp_b200.config.transformer_config",
    "megatron.core.ssm.mlp_layer": "megatronapp_b200.transformer.transformer_layer",
    "megatron.core.dist_checkpointing.optimizer": "megatronapp_b200.optim.optimizer",
    "megatron.core.dist_checkpointing.strategies.torch": "megatronapp_b200.ckpt.torch_dist",
    "megatron.core.dist_checkpointing.strategies.fully_parallel": "megatronapp_b200.ckpt.torch_dist",
    "megatron.core.dist_checkpointing.strategies.async_utils": "megatronapp_b200.training.async_utils",
    "megatron.core.dist_checkpointing.strategies.filesystem_async": "megatronapp_b200.training.async_utils",
    "megatron.core.dist_checkpointing.strategies.base": "megatronapp_b200.ckpt.serialization",
    "megatron.core.dist_checkpointing.strategies.common": "megatronapp_b200.ckpt.serialization",
    "megatron.core.dist_checkpointing.state_dict_utils": "megatronapp_b200.ckpt.serialization",
    "megatron.core.datasets.megatron_dataset": "megatronapp_b200.data.gpt_dataset",
    "megatron.core.datasets.utils": "megatronapp_b200.data.blended_builder",
    "megatron.core.datasets.multimodal_dataset": "megatronapp_b200.legacy.data.multimodal_dataset",
    "megatron.core.datasets.retro.utils": "megatronapp_b200.data.retro.config",
    "megatron.core.datasets.retro.query.gpt_chunk_dataset": "megatronapp_b200.data.retro.query",
    "megatron.core.datasets.retro.query.utils": "megatronapp_b200.data.retro.query",
    "megatron.core.datasets.retro.db.dataset": "megatronapp_b200.data.retro.db",
    "megatron.core.datasets.retro.db.utils": "megatronapp_b200.data.retro.db",
    "megatron.core.datasets.retro.index.build": "megatronapp_b200.data.retro.index",
    "megatron.core.datasets.retro.index.factory": "megatronapp_b200.data.retro.index",
    "megatron.core.datasets.retro.index.utils": "megatronapp_b200.data.retro.index",
    "megatron.core.export.trtllm.trt_model_config": "megatronapp_b200.export.trtllm",
    "megatron.core.export.trtllm.trtllm_layers": "megatronapp_b200.export.trtllm",
    "megatron.core.export.trtllm.model_to_trllm_mapping.default_conversion_dict": "megatronapp_b200.export.trtllm",
    "megatron.core.inference.contexts.base_context": "megatronapp_b200.inference.core",
    "megatron.core.inference.model_inference_wrappers.inference_wrapper_config": "megatronapp_b200.inference.core",
    "megatron.core.inference.utils": "megatronapp_b200.inference.core",
    "megatron.core.optimizer.cpu_offloading.hybrid_optimizer": "megatronapp_b200.optim.optimizer",
    "megatron.core.post_training.modelopt.layers": "megatronapp_b200.post_training.quantization",
    "megatron.core.post_training.modelopt.gpt.model_specs": "megatronapp_b200.models.gpt.gpt_layer_specs",
    "megatron.core.post_training.modelopt.gpt.state_dict_hooks": "megatronapp_b200.post_training.quantization",
    "megatron.core.post_training.modelopt.mamba.model_specs": "megatronapp_b200.models.mamba.mamba_model",
    "megatron.core.inference.modelopt_support.gpt.model_specs": "megatronapp_b200.models.gpt.gpt_layer_specs",
    "megatron.core.inference.modelopt_support.gpt.state_dict_hooks": "megatronapp_b200.post_training.quantization",
    "megatron.core.inference.modelopt_support.mamba.model_specs": "megatronapp_b200.models.mamba.mamba_model",
    "megatron.post_training.loss_func": "megatronapp_b200.post_training.distillation",
    "megatron.post_training.model_provider": "megatronapp_b200.post_training.distillation",
    "megatron.legacy.model.gpt_model": "megatronapp_b200.legacy.model",
    "megatron.legacy.model.t5_model": "megatronapp_b200.legacy.model",
    "megatron.legacy.model.language_model": "megatronapp_b200.legacy.model",
    "megatron.legacy.model.transformer": "megatronapp_b200.legacy.model",
    "megatron.legacy.model.realm_model": "megatronapp_b200.legacy.model.biencoder_model",
    "megatron.legacy.model.fused_layer_norm": "megatronapp_b200.transformer.norms",
    "megatron.legacy.model.rms_norm": "megatronapp_b200.transformer.norms",
    "megatron.legacy.model.vision.vit_backbone": "megatronapp_b200.models.vision.vit_classification",
    "megatron.legacy.data.dataset_utils": "megatronapp_b200.data.masked_dataset",
    "megatron.legacy.data.biencoder_dataset_utils": "megatronapp_b200.data.ict_dataset",
    "megatron.legacy.data.realm_dataset_utils": "megatronapp_b200.data.ict_dataset",
    "megatron.training.tokenizer.gpt2_tokenization": "megatronapp_b200.data.tokenizer",
    "megatron.training.tokenizer.bert_tokenization": "megatronapp_b200.data.tokenizer",
    "megatron.inference.text_generation.forward_step": "megatronapp_b200.inference.core",
    "megatron.inference.text_generation.mcore_engine_server": "megatronapp_b200.inference.server",
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
    "megatron.core.utils": ("megatronapp_b200.core_api.utils",),
    "megatron.core.transformer.utils": ("megatronapp_b200.core_api.transformer_utils", "megatronapp_b200.ckpt.sharded"),
    "megatron.legacy.model.utils": ("megatronapp_b200.core_api.transformer_utils", "megatronapp_b200.config.transformer_config",
                                    "megatronapp_b200.core_api.utils"),
    "megatron.core.dist_checkpointing.utils": ("megatronapp_b200.core_api.dist_ckpt_utils", "megatronapp_b200.ckpt.serialization",
                                               "megatronapp_b200.ckpt.sharded"),
    "megatron.core.dist_checkpointing.dict_utils": ("megatronapp_b200.core_api.dist_ckpt_utils",),
    "megatron.core.dist_checkpointing.mapping": ("megatronapp_b200.core_api.dist_ckpt_utils", "megatronapp_b200.ckpt",
                                                 "megatronapp_b200.ckpt.sharded", "megatronapp_b200.ckpt.serialization"),
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

        if (name in _ALIASES or any(k.startswith(name + ".") for k in _ALIASES) or name in _EXTRA_ATTRS or name in _UNIONS
                or any(k.startswith(name + ".") for k in _UNIONS)):
            return ModuleSpec(name, _Finder, is_package=name not in _ALIASES)
        return None

    @staticmethod
    def create_module(spec):
        if spec.name in _ALIASES:
            mod = importlib.import_module(_ALIASES[spec.name])
            if not hasattr(mod, "__path__") and any(k.startswith(spec.name + ".") for k in list(_ALIASES) + list(_UNIONS)):
                # the reference has a package where this framework has one module (``...common.embeddings`` and its
                # ``rotary_pos_embedding`` / ``language_model_embedding`` children): let the children be imported through it
                mod.__path__ = []
            return mod
        return _Lazy(spec.name)

    @staticmethod
    def exec_module(module):
        return None


if not any(isinstance(f, type) and f.__name__ == "_Finder" for f in sys.meta_path):
    sys.meta_path.insert(0, _Finder)
