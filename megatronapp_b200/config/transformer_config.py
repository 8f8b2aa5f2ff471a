"""Typed configuration tree.

``ModelParallelConfig`` / ``TransformerConfig`` keep the field names that user
code and the Megatron CLI rely on (reference ``model_parallel_config.py``,
``transformer/transformer_config.py:18-1125``), with validation in
``__post_init__``.  B200-specific knobs (fused TP kernels, fp8 recipe handled by
our own mxfp8 GEMMs) are additional fields.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple, Union

import torch
import torch.nn.functional as F


@dataclass
class ModelParallelConfig:
    tensor_model_parallel_size: int = 1
    pipeline_model_parallel_size: int = 1
    virtual_pipeline_model_parallel_size: Optional[int] = None
    sequence_parallel: bool = False
    context_parallel_size: int = 1
    hierarchical_context_parallel_sizes: Optional[List[int]] = None
    expert_model_parallel_size: int = 1
    expert_tensor_parallel_size: Optional[int] = None
    moe_extended_tp: bool = False

    perform_initialization: bool = True
    use_cpu_initialization: bool = False

    fp16: bool = False
    bf16: bool = False
    params_dtype: torch.dtype = torch.float32
    timers: Optional[object] = None
    finalize_model_grads_func: Optional[Callable] = None
    grad_scale_func: Optional[Callable] = None
    no_sync_func: Optional[Callable] = None
    grad_sync_func: Optional[Callable] = None
    param_sync_func: Optional[Callable] = None
    deterministic_mode: bool = False
    enable_autocast: bool = False
    autocast_dtype: Optional[torch.dtype] = None
    num_microbatches_with_partial_activation_checkpoints: Optional[int] = None

    gradient_accumulation_fusion: bool = False
    async_tensor_model_parallel_allreduce: bool = True
    use_te_rng_tracker: bool = False
    # fused compute+collective kernels over NVLink peer memory (the reference's
    # closest analogue is TE userbuffers, --tp-comm-overlap)
    tp_comm_overlap: bool = False
    tp_comm_bulk_wgrad: bool = True
    tp_comm_bulk_dgrad: bool = True
    tp_comm_overlap_ag: bool = True
    tp_comm_overlap_rs: bool = True
    tp_comm_overlap_rs_dgrad: bool = False
    tp_comm_split_ag: bool = True
    tp_comm_atomic_ag: bool = False
    tp_comm_split_rs: bool = True
    tp_comm_atomic_rs: bool = False
    cross_entropy_loss_fusion: bool = False
    tp_comm_overlap_disable_qkv: bool = False
    tp_comm_overlap_disable_fc1: bool = False
    tp_comm_bootstrap_backend: str = "nccl"

    pipeline_dtype: Optional[torch.dtype] = None
    variable_seq_lengths: bool = False
    overlap_p2p_comm: bool = False
    batch_p2p_comm: bool = True
    batch_p2p_sync: bool = True
    use_ring_exchange_p2p: bool = False
    deallocate_pipeline_outputs: bool = False
    defer_embedding_wgrad_compute: bool = False
    wgrad_deferral_limit: int = 0
    pipeline_model_parallel_split_rank: Optional[int] = None
    overlap_p2p_comm_warmup_flush: bool = False
    microbatch_group_size_per_vp_stage: Optional[int] = None

    cpu_offloading: bool = False
    cpu_offloading_num_layers: int = 0
    cpu_offloading_activations: bool = True
    cpu_offloading_weights: bool = True
    barrier_with_L1_time: bool = True

    def __post_init__(self):
        if self.sequence_parallel and self.tensor_model_parallel_size <= 1:
            # sequence parallel is a no-op at tp == 1; keep reference behaviour (silently off)
            self.sequence_parallel = False
        if self.expert_tensor_parallel_size is None:
            self.expert_tensor_parallel_size = self.tensor_model_parallel_size
        if self.pipeline_model_parallel_size > 1 and self.pipeline_dtype is None:
            self.pipeline_dtype = self.params_dtype
        if self.autocast_dtype is None:
            self.autocast_dtype = self.params_dtype
        if self.defer_embedding_wgrad_compute and self.pipeline_model_parallel_size == 1:
            raise ValueError("defer_embedding_wgrad_compute needs pipeline parallelism")
        if self.microbatch_group_size_per_vp_stage is None:
            self.microbatch_group_size_per_vp_stage = self.pipeline_model_parallel_size


def _gelu(x):
    return F.gelu(x)


@dataclass
class TransformerConfig(ModelParallelConfig):
    num_layers: int = 0
    mtp_num_layers: Optional[int] = None
    mtp_loss_scaling_factor: Optional[float] = None
    num_layers_in_first_pipeline_stage: Optional[int] = None
    num_layers_in_last_pipeline_stage: Optional[int] = None
    account_for_embedding_in_pipeline_split: bool = False
    account_for_loss_in_pipeline_split: bool = False
    hidden_size: int = 0
    num_attention_heads: int = 0
    attention_backend: str = "auto"
    softmax_scale: Optional[float] = None
    num_query_groups: Optional[int] = None
    ffn_hidden_size: Optional[int] = None
    kv_channels: Optional[int] = None
    hidden_dropout: float = 0.1
    attention_dropout: float = 0.1
    fp32_residual_connection: bool = False
    apply_residual_connection_post_layernorm: bool = False
    layernorm_epsilon: float = 1e-5
    layernorm_zero_centered_gamma: bool = False
    add_bias_linear: bool = True
    add_qkv_bias: bool = False
    gated_linear_unit: bool = False
    activation_func: Callable = _gelu
    activation_func_fp8_input_store: bool = False
    num_moe_experts: Optional[int] = None
    rotary_interleaved: bool = False
    window_size: Optional[Tuple[int, int]] = None
    normalization: str = "LayerNorm"
    qk_layernorm: bool = False
    test_mode: bool = False
    calculate_per_token_loss: bool = False
    multi_latent_attention: bool = False

    init_method: Optional[Callable] = None
    output_layer_init_method: Optional[Callable] = None
    init_method_std: float = 0.02
    init_model_with_meta_device: bool = False

    apply_query_key_layer_scaling: bool = False
    attention_softmax_in_fp32: bool = True
    disable_bf16_reduced_precision_matmul: bool = False

    bias_activation_fusion: bool = False
    masked_softmax_fusion: bool = False
    persist_layer_norm: bool = False
    memory_efficient_layer_norm: bool = False
    bias_dropout_fusion: bool = False
    apply_rope_fusion: bool = False

    recompute_granularity: Optional[str] = None
    recompute_method: Optional[str] = None
    recompute_num_layers: Optional[int] = None
    distribute_saved_activations: Optional[bool] = None
    recompute_modules: Optional[List[str]] = None

    mrope_section: Optional[list] = None           # M-RoPE: rotary channel pairs driven by (temporal, height, width) positions
    fbd_backward_config: Optional[object] = None   # collapsed MegaFBD: the (tp, SP) layout of this forward rank's backward twins
    fp8: Optional[str] = None
    fp8_recipe: Optional[str] = "delayed"
    fp8_param: bool = False
    fp8_margin: int = 0
    fp8_interval: int = 1
    fp8_amax_history_len: int = 1
    fp8_amax_compute_algo: str = "most_recent"
    fp8_wgrad: bool = True
    fp8_dot_product_attention: bool = False
    fp8_multi_head_attention: bool = False
    tp_only_amax_red: bool = False
    first_last_layers_bf16: bool = False
    num_layers_at_start_in_bf16: int = 1
    num_layers_at_end_in_bf16: int = 1

    moe_shared_expert_intermediate_size: Optional[int] = None
    moe_shared_expert_overlap: bool = False
    moe_shared_expert_gate: bool = False
    moe_layer_freq: Union[int, List[int]] = 1
    moe_ffn_hidden_size: Optional[int] = None
    moe_router_load_balancing_type: str = "aux_loss"
    moe_router_topk: int = 2
    moe_router_topk_limited_devices: Optional[int] = None
    moe_router_num_groups: Optional[int] = None
    moe_router_group_topk: Optional[int] = None
    moe_router_pre_softmax: bool = False
    moe_router_topk_scaling_factor: Optional[float] = None
    moe_router_score_function: str = "softmax"
    moe_router_dtype: Optional[str] = None
    moe_router_enable_expert_bias: bool = False
    moe_router_bias_update_rate: float = 1e-3
    moe_grouped_gemm: bool = False
    moe_use_legacy_grouped_gemm: bool = False
    moe_aux_loss_coeff: float = 0.0
    moe_z_loss_coeff: Optional[float] = None
    moe_input_jitter_eps: Optional[float] = None
    moe_token_dropping: bool = False
    moe_token_dispatcher_type: str = "allgather"
    moe_enable_deepep: bool = False
    moe_per_layer_logging: bool = False
    moe_expert_capacity_factor: Optional[float] = None
    moe_pad_expert_input_to_capacity: bool = False
    moe_token_drop_policy: str = "probs"
    moe_layer_recompute: bool = False
    moe_permute_fusion: bool = False
    moe_apply_probs_on_input: bool = False

    cp_comm_type: Optional[Union[str, List[str]]] = None
    enable_cuda_graph: bool = False
    cuda_graph_use_single_mempool: bool = False
    cuda_graph_retain_backward_graph: bool = False
    cuda_graph_warmup_steps: int = 3
    external_cuda_graph: bool = False
    cuda_graph_scope: str = "full"
    clone_scatter_output_in_embedding: bool = True
    disable_parameter_transpose_cache: bool = False
    config_logger_dir: str = ""
    flash_decode: bool = False
    inference_rng_tracker: bool = False
    use_custom_fsdp: bool = False
    is_hybrid_model: bool = False
    mamba_state_dim: int = 128
    mamba_head_dim: int = 64
    mamba_num_groups: int = 8
    heterogeneous_block_specs: bool = False

    # --- B200-native additions ------------------------------------------
    # "b200": hand-written sm_100a kernels (default on CUDA); "torch": PyTorch
    # reference ops (CPU / debugging).  --transformer-impl local maps to "b200"
    # on a GPU: the add-ons work on the fast path (no local/TE split).
    kernel_backend: str = "auto"
    # MegaScope hooks are active only when a capture registry is installed
    scope_enabled: bool = False

    def __post_init__(self):
        super().__post_init__()
        if self.fp16 and self.bf16:
            raise ValueError("only one of fp16 and bf16 may be set")
        if self.num_attention_heads % max(1, self.tensor_model_parallel_size) != 0:
            raise ValueError(
                f"num_attention_heads ({self.num_attention_heads}) must be a multiple of "
                f"tensor_model_parallel_size ({self.tensor_model_parallel_size})")
        if self.ffn_hidden_size is None:
            self.ffn_hidden_size = 4 * self.hidden_size
        if self.kv_channels is None:
            self.kv_channels = self.hidden_size // self.num_attention_heads
        if self.num_query_groups is None:
            self.num_query_groups = self.num_attention_heads
        if self.num_query_groups % max(1, self.tensor_model_parallel_size) != 0:
            raise ValueError(
                f"num_query_groups ({self.num_query_groups}) must be a multiple of "
                f"tensor_model_parallel_size ({self.tensor_model_parallel_size})")
        if self.apply_query_key_layer_scaling:
            self.attention_softmax_in_fp32 = True
        if self.expert_model_parallel_size > 1 and self.num_moe_experts is None:
            raise ValueError("num_moe_experts must be set for expert parallelism")
        if self.num_moe_experts is not None and self.num_moe_experts <= 0:
            raise ValueError("num_moe_experts must be positive")
        if (self.num_moe_experts is not None
                and self.num_moe_experts % self.expert_model_parallel_size != 0):
            raise ValueError("num_moe_experts must be divisible by expert_model_parallel_size")
        if self.moe_ffn_hidden_size is None:
            self.moe_ffn_hidden_size = self.ffn_hidden_size
        if self.moe_expert_capacity_factor is not None:
            if self.moe_expert_capacity_factor < 0:
                self.moe_expert_capacity_factor = None
            if self.moe_router_load_balancing_type not in ("aux_loss", "seq_aux_loss", "none"):
                raise ValueError("capacity factor needs aux_loss / seq_aux_loss / none balancing")
        if self.moe_pad_expert_input_to_capacity and self.moe_expert_capacity_factor is None:
            raise ValueError("moe_pad_expert_input_to_capacity needs moe_expert_capacity_factor")
        if self.cpu_offloading and self.cpu_offloading_num_layers not in range(0, max(1, self.num_layers)):
            raise ValueError("cpu_offloading_num_layers out of range")
        if self.recompute_granularity is not None:
            if self.recompute_granularity not in ("full", "selective"):
                raise ValueError(f"bad recompute_granularity {self.recompute_granularity}")
            if self.recompute_granularity == "full":
                if self.recompute_method not in ("block", "uniform"):
                    raise ValueError('recompute_method must be "block" or "uniform" with full recompute')
                if self.recompute_num_layers is None:
                    raise ValueError("recompute_num_layers must be set with full recompute")
            if self.recompute_granularity == "selective" and self.recompute_modules is None:
                self.recompute_modules = ["core_attn"]
        if self.distribute_saved_activations and self.sequence_parallel:
            raise ValueError("distribute_saved_activations and sequence_parallel are exclusive")
        if self.virtual_pipeline_model_parallel_size is not None:
            pp = self.pipeline_model_parallel_size
            if (self.num_layers // pp) % self.virtual_pipeline_model_parallel_size != 0:
                raise ValueError("layers per pipeline stage must divide evenly into virtual stages")
        if self.bias_activation_fusion and self.activation_func not in (F.gelu, F.silu, _gelu):
            raise ValueError("bias_activation_fusion supports gelu / silu(swiglu)")
        if self.init_method is None:
            self.init_method = init_method_normal(self.init_method_std)
        if self.output_layer_init_method is None:
            self.output_layer_init_method = scaled_init_method_normal(
                self.init_method_std, max(1, self.num_layers))
        if self.fp8 is not None and self.fp8 not in ("e4m3", "hybrid"):
            raise ValueError("fp8 format must be e4m3 or hybrid")
        if isinstance(self.moe_layer_freq, list) and len(self.moe_layer_freq) != self.num_layers:
            raise ValueError("moe_layer_freq list must have num_layers entries")
        if self.cp_comm_type is not None and isinstance(self.cp_comm_type, list):
            if len(self.cp_comm_type) != self.num_layers:
                raise ValueError("cp_comm_type list must have num_layers entries")
        if self.kernel_backend not in ("auto", "b200", "torch"):
            raise ValueError("kernel_backend must be auto | b200 | torch")

    @property
    def use_b200_kernels(self) -> bool:
        if self.kernel_backend == "torch":
            return False
        return torch.cuda.is_available()


@dataclass
class MLATransformerConfig(TransformerConfig):
    """DeepSeek multi-latent attention (reference ``transformer_config.py:1072``)."""

    multi_latent_attention: bool = True
    q_lora_rank: Optional[int] = 512
    kv_lora_rank: int = 512
    qk_head_dim: int = 128
    qk_pos_emb_head_dim: int = 64
    v_head_dim: int = 128
    normalization: str = "RMSNorm"
    rope_type: str = "yarn"
    rotary_base: float = 10000
    rotary_percent: float = 1.0
    rotary_scaling_factor: float = 40
    max_position_embeddings: int = 4096
    beta_fast: float = 32
    beta_slow: float = 1
    mscale: float = 0.707
    mscale_all_dim: float = 0.707


def init_method_normal(sigma: float):
    def init_(tensor):
        return torch.nn.init.normal_(tensor, mean=0.0, std=sigma)

    return init_


def scaled_init_method_normal(sigma: float, num_layers: int):
    std = sigma / (2.0 * num_layers) ** 0.5

    def init_(tensor):
        return torch.nn.init.normal_(tensor, mean=0.0, std=std)

    return init_
