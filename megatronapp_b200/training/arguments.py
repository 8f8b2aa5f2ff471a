"""Megatron-compatible command line (flag names kept) -> validated namespace -> config objects.

Parity: reference ``megatron/training/arguments.py`` (``parse_args`` :70, ``validate_args``
:286-983, ``core_transformer_config_from_args`` :1004-1057, add-on flags :1663 (``--training-ws-port``),
:2197-2205 (FBD / DPP), :2705-2717 (MegaScan)).  Flags that only select between library back-ends
in the reference (``--transformer-impl``, ``--attention-backend``, ``--use-flash-attn``) are
accepted and mapped onto the single native path.
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
from typing import Callable, Optional

import torch
import torch.nn.functional as F

from ..config.transformer_config import MLATransformerConfig, TransformerConfig


def _add(group, *names, **kw):
    group.add_argument(*names, **kw)


def build_parser(extra_args_provider: Optional[Callable] = None) -> argparse.ArgumentParser:
    # "resolve": an entry script may re-declare a flag (with its own default) without tripping over the central one
    p = argparse.ArgumentParser(description="megatronapp_b200 arguments", allow_abbrev=False, conflict_handler="resolve")
    g = p.add_argument_group("network size")
    _add(g, "--num-layers", type=int, default=None)
    _add(g, "--encoder-num-layers", type=int, default=None)
    _add(g, "--decoder-num-layers", type=int, default=None)
    _add(g, "--hidden-size", type=int, default=None)
    _add(g, "--ffn-hidden-size", type=int, default=None)
    _add(g, "--num-attention-heads", type=int, default=None)
    _add(g, "--kv-channels", type=int, default=None)
    _add(g, "--group-query-attention", action="store_true")
    _add(g, "--num-query-groups", type=int, default=1)
    _add(g, "--max-position-embeddings", type=int, default=None)
    _add(g, "--position-embedding-type", type=str, default="learned_absolute",
         choices=["learned_absolute", "rope", "relative", "none", "mrope"])
    _add(g, "--use-rotary-position-embeddings", action="store_true")
    _add(g, "--rotary-base", type=int, default=10000)
    _add(g, "--rotary-percent", type=float, default=1.0)
    _add(g, "--rotary-interleaved", action="store_true")
    _add(g, "--rotary-seq-len-interpolation-factor", type=int, default=None)
    _add(g, "--use-rope-scaling", action="store_true")
    _add(g, "--rope-scaling-factor", type=float, default=8.0)
    _add(g, "--no-position-embedding", action="store_false", dest="add_position_embedding")
    _add(g, "--make-vocab-size-divisible-by", type=int, default=128)
    _add(g, "--normalization", default="LayerNorm", choices=["LayerNorm", "RMSNorm"])
    _add(g, "--norm-epsilon", type=float, default=1e-5)
    _add(g, "--layernorm-epsilon", type=float, default=None)
    _add(g, "--apply-layernorm-1p", action="store_true")
    _add(g, "--apply-residual-connection-post-layernorm", action="store_true")
    _add(g, "--openai-gelu", action="store_true")
    _add(g, "--squared-relu", action="store_true")
    _add(g, "--swiglu", action="store_true")
    _add(g, "--onnx-safe", type=bool, default=None)
    _add(g, "--bert-no-binary-head", action="store_false", dest="bert_binary_head")
    _add(g, "--untie-embeddings-and-output-weights", action="store_true")
    _add(g, "--multi-latent-attention", action="store_true")
    # reference _add_mla_args: low-rank q / kv projections, decoupled rope head, YaRN scale factors
    _add(g, "--q-lora-rank", type=int, default=None)
    _add(g, "--kv-lora-rank", type=int, default=32)
    _add(g, "--qk-head-dim", type=int, default=128)
    _add(g, "--qk-pos-emb-head-dim", type=int, default=64)
    _add(g, "--v-head-dim", type=int, default=128)
    _add(g, "--rotary-scaling-factor", type=float, default=1.0)
    _add(g, "--mscale", type=float, default=1.0)
    _add(g, "--mscale-all-dim", type=float, default=1.0)
    _add(g, "--qk-layernorm", action="store_true")
    _add(g, "--disable-bias-linear", action="store_false", dest="add_bias_linear")
    _add(g, "--add-qkv-bias", action="store_true")
    _add(g, "--mtp-num-layers", type=int, default=None)
    _add(g, "--mtp-loss-scaling-factor", type=float, default=0.1)
    _add(g, "--hybrid-attention-ratio", type=float, default=0.0)
    _add(g, "--hybrid-mlp-ratio", type=float, default=0.0)
    _add(g, "--hybrid-override-pattern", type=str, default=None)
    _add(g, "--mamba-state-dim", type=int, default=128)
    _add(g, "--mamba-head-dim", type=int, default=64)
    _add(g, "--mamba-num-groups", type=int, default=8)
    _add(g, "--is-hybrid-model", action="store_true")
    g = p.add_argument_group("moe")
    _add(g, "--num-experts", type=int, default=None)
    _add(g, "--expert-model-parallel-size", type=int, default=1)
    _add(g, "--expert-tensor-parallel-size", type=int, default=None)
    _add(g, "--moe-layer-freq", type=str, default="1")
    _add(g, "--moe-ffn-hidden-size", type=int, default=None)
    _add(g, "--moe-shared-expert-intermediate-size", type=int, default=None)
    _add(g, "--moe-shared-expert-overlap", action="store_true")
    _add(g, "--moe-shared-expert-gate", action="store_true", help="scale the shared expert by a learned sigmoid gate (Qwen2-MoE)")
    _add(g, "--moe-grouped-gemm", action="store_true")
    _add(g, "--moe-use-legacy-grouped-gemm", action="store_true")
    _add(g, "--moe-router-load-balancing-type", type=str, default="aux_loss",
         choices=["aux_loss", "seq_aux_loss", "sinkhorn", "none"])
    _add(g, "--moe-router-dtype", type=str, default=None, choices=["fp32", "fp64"])
    _add(g, "--moe-router-score-function", type=str, default="softmax", choices=["softmax", "sigmoid"])
    _add(g, "--moe-router-topk", type=int, default=2)
    _add(g, "--moe-router-pre-softmax", action="store_true")
    _add(g, "--moe-router-num-groups", type=int, default=None)
    _add(g, "--moe-router-group-topk", type=int, default=None)
    _add(g, "--moe-router-topk-scaling-factor", type=float, default=None)
    _add(g, "--moe-router-enable-expert-bias", action="store_true")
    _add(g, "--moe-router-bias-update-rate", type=float, default=1e-3)
    _add(g, "--moe-aux-loss-coeff", type=float, default=0.0)
    _add(g, "--moe-z-loss-coeff", type=float, default=None)
    _add(g, "--moe-input-jitter-eps", type=float, default=None)
    _add(g, "--moe-token-dispatcher-type", type=str, default="allgather",
         choices=["allgather", "alltoall", "flex", "alltoall_seq"])
    _add(g, "--moe-enable-deepep", action="store_true")
    _add(g, "--moe-per-layer-logging", action="store_true")
    _add(g, "--moe-expert-capacity-factor", type=float, default=None)
    _add(g, "--moe-pad-expert-input-to-capacity", action="store_true")
    _add(g, "--moe-token-drop-policy", type=str, default="probs", choices=["probs", "position"])
    _add(g, "--moe-layer-recompute", action="store_true")
    _add(g, "--moe-extended-tp", action="store_true")
    _add(g, "--moe-permute-fusion", action="store_true")
    _add(g, "--moe-use-upcycling", action="store_true")
    g = p.add_argument_group("regularization")
    _add(g, "--attention-dropout", type=float, default=0.1)
    _add(g, "--hidden-dropout", type=float, default=0.1)
    _add(g, "--weight-decay", type=float, default=0.01)
    _add(g, "--start-weight-decay", type=float, default=None)
    _add(g, "--end-weight-decay", type=float, default=None)
    _add(g, "--weight-decay-incr-style", type=str, default="constant", choices=["constant", "linear", "cosine"])
    _add(g, "--clip-grad", type=float, default=1.0)
    _add(g, "--adam-beta1", type=float, default=0.9)
    _add(g, "--adam-beta2", type=float, default=0.999)
    _add(g, "--adam-eps", type=float, default=1e-8)
    _add(g, "--sgd-momentum", type=float, default=0.9)
    g = p.add_argument_group("training")
    _add(g, "--micro-batch-size", type=int, default=None)
    _add(g, "--batch-size", type=int, default=None)
    _add(g, "--global-batch-size", type=int, default=None)
    _add(g, "--rampup-batch-size", nargs="*", default=None)
    _add(g, "--decrease-batch-size-if-needed", action="store_true")
    _add(g, "--recompute-activations", action="store_true")
    _add(g, "--recompute-granularity", type=str, default=None, choices=["full", "selective"])
    _add(g, "--recompute-method", type=str, default=None, choices=["uniform", "block"])
    _add(g, "--recompute-num-layers", type=int, default=None)
    _add(g, "--recompute-modules", nargs="*", type=str, default=None)
    _add(g, "--distribute-saved-activations", action="store_true")
    _add(g, "--no-check-for-nan-in-loss-and-grad", action="store_false", dest="check_for_nan_in_loss_and_grad")
    _add(g, "--check-for-spiky-loss", action="store_true")
    _add(g, "--check-for-large-grads", action="store_true")
    _add(g, "--deterministic-mode", action="store_true")
    _add(g, "--check-weight-hash-across-dp-replicas-interval", type=int, default=None)
    _add(g, "--calculate-per-token-loss", action="store_true")
    _add(g, "--train-iters", type=int, default=None)
    _add(g, "--train-samples", type=int, default=None)
    _add(g, "--log-interval", type=int, default=100)
    _add(g, "--exit-interval", type=int, default=None)
    _add(g, "--exit-duration-in-mins", type=int, default=None)
    _add(g, "--exit-signal-handler", action="store_true")
    _add(g, "--adlr-autoresume", action="store_true")
    _add(g, "--adlr-autoresume-interval", type=int, default=1000)
    _add(g, "--tensorboard-dir", type=str, default=None)
    _add(g, "--no-masked-softmax-fusion", action="store_false", dest="masked_softmax_fusion")
    _add(g, "--no-bias-gelu-fusion", action="store_false", dest="bias_gelu_fusion")
    _add(g, "--no-bias-swiglu-fusion", action="store_false", dest="bias_swiglu_fusion")
    _add(g, "--no-bias-dropout-fusion", action="store_false", dest="bias_dropout_fusion")
    _add(g, "--no-rope-fusion", action="store_false", dest="apply_rope_fusion")
    _add(g, "--cross-entropy-loss-fusion", action="store_true")
    _add(g, "--use-flash-attn", action="store_true")
    _add(g, "--attention-backend", type=str, default="auto")
    _add(g, "--optimizer", type=str, default="adam", choices=["adam", "sgd"])
    _add(g, "--dataloader-type", type=str, default=None, choices=["single", "cyclic", "external"])
    _add(g, "--no-async-tensor-model-parallel-allreduce", action="store_false", dest="async_tensor_model_parallel_allreduce")
    _add(g, "--no-persist-layer-norm", action="store_true")
    _add(g, "--sequence-parallel", action="store_true")
    _add(g, "--no-gradient-accumulation-fusion", action="store_false", dest="gradient_accumulation_fusion")
    _add(g, "--use-mcore-models", action="store_true")
    _add(g, "--use-legacy-models", action="store_true")
    _add(g, "--manual-gc", action="store_true")
    _add(g, "--manual-gc-interval", type=int, default=0)
    _add(g, "--no-manual-gc-eval", action="store_false", dest="manual_gc_eval")
    _add(g, "--disable-tp-comm-split-ag", action="store_false", dest="tp_comm_split_ag")
    _add(g, "--disable-tp-comm-split-rs", action="store_false", dest="tp_comm_split_rs")
    _add(g, "--tp-comm-overlap", action="store_true")
    _add(g, "--tp-comm-overlap-cfg", type=str, default=None)
    _add(g, "--disable-tp-comm-overlap-ag", action="store_false", dest="tp_comm_overlap_ag")
    _add(g, "--disable-tp-comm-overlap-rs", action="store_false", dest="tp_comm_overlap_rs")
    _add(g, "--tp-comm-overlap-rs-dgrad", action="store_true")
    _add(g, "--disable-tp-comm-bulk-dgrad", action="store_false", dest="tp_comm_bulk_dgrad")
    _add(g, "--disable-tp-comm-bulk-wgrad", action="store_false", dest="tp_comm_bulk_wgrad")
    _add(g, "--tp-comm-bootstrap-backend", default="nccl", type=str)
    _add(g, "--use-cpu-initialization", action="store_true", default=None)
    _add(g, "--empty-unused-memory-level", default=0, type=int, choices=[0, 1, 2])
    _add(g, "--iterations-to-skip", nargs="+", type=int, default=[])
    _add(g, "--profile", action="store_true")
    _add(g, "--profile-step-start", type=int, default=10)
    _add(g, "--profile-step-end", type=int, default=12)
    _add(g, "--profile-ranks", nargs="+", type=int, default=[0])
    _add(g, "--use-pytorch-profiler", action="store_true")
    _add(g, "--record-memory-history", action="store_true")
    _add(g, "--memory-snapshot-path", type=str, default="snapshot.pickle")
    _add(g, "--cuda-graph-warmup-steps", type=int, default=3)
    _add(g, "--enable-cuda-graph", action="store_true")
    _add(g, "--kernel-backend", type=str, default="auto", choices=["auto", "b200", "torch"])
    _add(g, "--no-fused-tp-comm", action="store_false", dest="fused_tp_comm",
         help="use NCCL collectives instead of the fused AG/RS GEMM kernels (baseline mode)")
    g = p.add_argument_group("initialization")
    _add(g, "--seed", type=int, default=1234)
    _add(g, "--data-parallel-random-init", action="store_true")
    _add(g, "--init-method-std", type=float, default=0.02)
    _add(g, "--init-method-xavier-uniform", action="store_true")
    g = p.add_argument_group("learning rate")
    _add(g, "--lr", type=float, default=None)
    _add(g, "--lr-decay-style", type=str, default="linear",
         choices=["constant", "linear", "cosine", "inverse-square-root", "WSD"])
    _add(g, "--lr-wsd-decay-style", type=str, default="exponential")
    _add(g, "--lr-decay-iters", type=int, default=None)
    _add(g, "--lr-decay-samples", type=int, default=None)
    _add(g, "--lr-wsd-decay-iters", type=int, default=None)
    _add(g, "--lr-wsd-decay-samples", type=int, default=None)
    _add(g, "--lr-warmup-fraction", type=float, default=None)
    _add(g, "--lr-warmup-iters", type=int, default=0)
    _add(g, "--lr-warmup-samples", type=int, default=0)
    _add(g, "--lr-warmup-init", type=float, default=0.0)
    _add(g, "--min-lr", type=float, default=0.0)
    _add(g, "--override-opt_param-scheduler", action="store_true")
    _add(g, "--use-checkpoint-opt_param-scheduler", action="store_true")
    _add(g, "--decoupled-lr", type=float, default=None)
    _add(g, "--decoupled-min-lr", type=float, default=None)
    g = p.add_argument_group("checkpointing")
    _add(g, "--save", type=str, default=None)
    _add(g, "--save-interval", "--persistent-save-interval", type=int, default=None)
    _add(g, "--no-save-optim", action="store_true", default=None)
    _add(g, "--no-save-rng", action="store_true", default=None)
    _add(g, "--load", type=str, default=None)
    _add(g, "--no-load-optim", action="store_true", default=None)
    _add(g, "--no-load-rng", action="store_true", default=None)
    _add(g, "--non-persistent-save-interval", type=int, default=None)
    _add(g, "--non-persistent-ckpt-type", type=str, default=None, choices=["global", "local", "in_memory", None])
    _add(g, "--non-persistent-global-ckpt-dir", type=str, default=None)
    _add(g, "--finetune", action="store_true")
    _add(g, "--pretrained-checkpoint", type=str, default=None)
    _add(g, "--ckpt-step", type=int, default=None)
    _add(g, "--no-initialization", action="store_false", dest="perform_initialization")
    _add(g, "--use-checkpoint-args", action="store_true")
    _add(g, "--exit-on-missing-checkpoint", action="store_true")
    _add(g, "--use-dist-ckpt", action="store_true", dest="use_dist_ckpt_deprecated")
    _add(g, "--auto-detect-ckpt-format", action="store_true")
    _add(g, "--dist-ckpt-format", dest="dist_ckpt_format_deprecated", default=None)
    _add(g, "--ckpt-format", default="torch_dist", choices=["torch", "torch_dist", "zarr", "torch_dcp", "b200_sharded"],
         help="zarr / torch_dcp are accepted for compatibility and written in the torch_dist layout")
    _add(g, "--ckpt-convert-format", default=None, choices=["torch", "torch_dist", "zarr"])
    _add(g, "--ckpt-convert-save", default=None)
    _add(g, "--ckpt-fully-parallel-save", action="store_true", dest="ckpt_fully_parallel_save_deprecated")
    _add(g, "--no-ckpt-fully-parallel-save", action="store_false", dest="ckpt_fully_parallel_save")
    _add(g, "--async-save", action="store_true", default=None)
    _add(g, "--ckpt-fully-parallel-load", action="store_true")
    _add(g, "--ckpt-assume-constant-structure", action="store_true")
    _add(g, "--dist-ckpt-strictness", type=str, default="assume_ok_unexpected",
         choices=["assume_ok_unexpected", "log_unexpected", "log_all", "raise_unexpected", "raise_all", "return_unexpected",
                  "return_all", "ignore_all"])
    g = p.add_argument_group("mixed precision")
    _add(g, "--fp16", action="store_true")
    _add(g, "--bf16", action="store_true")
    _add(g, "--loss-scale", type=float, default=None)
    _add(g, "--initial-loss-scale", type=float, default=2 ** 32)
    _add(g, "--min-loss-scale", type=float, default=1.0)
    _add(g, "--loss-scale-window", type=float, default=1000)
    _add(g, "--hysteresis", type=int, default=2)
    _add(g, "--fp32-residual-connection", action="store_true")
    _add(g, "--apply-query-key-layer-scaling", action="store_true")
    _add(g, "--attention-softmax-in-fp32", action="store_true")
    _add(g, "--accumulate-allreduce-grads-in-fp32", action="store_true")
    _add(g, "--fp16-lm-cross-entropy", action="store_true")
    _add(g, "--fp8-format", default=None, choices=["e4m3", "hybrid"], dest="fp8")
    _add(g, "--fp8-recipe", default="delayed", choices=["tensorwise", "delayed", "mxfp8"])
    _add(g, "--fp8-margin", type=int, default=0)
    _add(g, "--fp8-interval", type=int, default=1)
    _add(g, "--fp8-amax-history-len", type=int, default=1)
    _add(g, "--fp8-amax-compute-algo", default="most_recent", choices=["most_recent", "max"])
    _add(g, "--no-fp8-wgrad", action="store_false", dest="fp8_wgrad")
    _add(g, "--fp8-param-gather", action="store_true")
    _add(g, "--first-last-layers-bf16", action="store_true")
    _add(g, "--num-layers-at-start-in-bf16", type=int, default=1)
    _add(g, "--num-layers-at-end-in-bf16", type=int, default=1)
    _add(g, "--transformer-impl", default="local", choices=["local", "transformer_engine"])
    g = p.add_argument_group("distributed")
    _add(g, "--tensor-model-parallel-size", type=int, default=1)
    _add(g, "--pipeline-model-parallel-size", type=int, default=1)
    _add(g, "--encoder-tensor-model-parallel-size", type=int, default=0)
    _add(g, "--encoder-pipeline-model-parallel-size", type=int, default=0)
    _add(g, "--pipeline-model-parallel-split-rank", type=int, default=None)
    _add(g, "--decoder-first-pipeline-num-layers", type=int, default=None)
    _add(g, "--decoder-last-pipeline-num-layers", type=int, default=None)
    _add(g, "--model-parallel-size", type=int, default=None)
    _add(g, "--num-layers-per-virtual-pipeline-stage", type=int, default=None)
    _add(g, "--num-virtual-stages-per-pipeline-rank", type=int, default=None)
    _add(g, "--microbatch-group-size-per-virtual-pipeline-stage", type=int, default=None)
    _add(g, "--no-overlap-p2p-communication", action="store_false", dest="overlap_p2p_comm")
    _add(g, "--overlap-p2p-communication-warmup-flush", action="store_true", dest="overlap_p2p_comm_warmup_flush")
    _add(g, "--distributed-backend", default="nccl", choices=["nccl", "gloo"])
    _add(g, "--distributed-timeout-minutes", type=int, default=10)
    _add(g, "--overlap-grad-reduce", action="store_true")
    _add(g, "--defer-embedding-wgrad-compute", action="store_true")
    _add(g, "--wgrad-deferral-limit", type=int, default=0)
    _add(g, "--no-align-grad-reduce", action="store_false", dest="align_grad_reduce")
    _add(g, "--ddp-num-buckets", type=int, default=None)
    _add(g, "--ddp-bucket-size", type=int, default=None)
    _add(g, "--ddp-average-in-collective", action="store_true")
    _add(g, "--overlap-param-gather", action="store_true")
    _add(g, "--overlap-param-gather-with-optimizer-step", action="store_true")
    _add(g, "--no-align-param-gather", action="store_false", dest="align_param_gather")
    _add(g, "--no-scatter-gather-tensors-in-pipeline", action="store_false", dest="scatter_gather_tensors_in_pipeline")
    _add(g, "--use-ring-exchange-p2p", action="store_true")
    _add(g, "--local-rank", "--local_rank", type=int, default=int(os.getenv("LOCAL_RANK", "0")))
    _add(g, "--lazy-mpu-init", type=bool, required=False)
    _add(g, "--account-for-embedding-in-pipeline-split", action="store_true")
    _add(g, "--account-for-loss-in-pipeline-split", action="store_true")
    _add(g, "--use-distributed-optimizer", action="store_true")
    _add(g, "--num-distributed-optimizer-instances", type=int, default=1)
    _add(g, "--use-torch-fsdp2", action="store_true")
    _add(g, "--use-custom-fsdp", action="store_true")
    _add(g, "--data-parallel-sharding-strategy", type=str, default="no_shard",
         choices=["no_shard", "optim", "optim_grads", "optim_grads_params"])
    _add(g, "--context-parallel-size", type=int, default=1)
    _add(g, "--cp-comm-type", nargs="+", type=str, default=["p2p"])
    _add(g, "--hierarchical-context-parallel-sizes", nargs="+", type=int, default=None)
    _add(g, "--nccl-communicator-config-path", type=str, default=None)
    _add(g, "--use-tp-pp-dp-mapping", action="store_true")
    _add(g, "--replication", action="store_true")
    _add(g, "--fused-grad-reduce-scatter", action="store_true",
         help="DP gradient reduce-scatter as one sm_100a kernel fused with scale/cast over NVLink")
    g = p.add_argument_group("validation")
    _add(g, "--eval-iters", type=int, default=100)
    _add(g, "--eval-interval", type=int, default=1000)
    _add(g, "--test-mode", action="store_true")
    _add(g, "--skip-train", action="store_true")
    g = p.add_argument_group("data and dataloader")
    _add(g, "--data-path", nargs="*", default=None)
    _add(g, "--split", type=str, default=None)
    _add(g, "--train-data-path", nargs="*", default=None)
    _add(g, "--valid-data-path", nargs="*", default=None)
    _add(g, "--test-data-path", nargs="*", default=None)
    _add(g, "--data-args-path", type=str, default=None)
    _add(g, "--per-split-data-args-path", type=str, default=None)
    _add(g, "--data-cache-path", default=None)
    _add(g, "--no-mmap-bin-files", action="store_false", dest="mmap_bin_files")
    _add(g, "--mock-data", action="store_true")
    _add(g, "--seq-length", type=int, default=None)
    _add(g, "--encoder-seq-length", type=int, default=None)
    _add(g, "--decoder-seq-length", type=int, default=None)
    _add(g, "--retriever-seq-length", type=int, default=256)
    _add(g, "--sample-rate", type=float, default=1.0)
    _add(g, "--mask-prob", type=float, default=0.15)
    _add(g, "--short-seq-prob", type=float, default=0.1)
    _add(g, "--num-workers", type=int, default=2)
    _add(g, "--reset-position-ids", action="store_true")
    _add(g, "--reset-attention-mask", action="store_true")
    _add(g, "--eod-mask-loss", action="store_true")
    _add(g, "--no-create-attention-mask-in-dataloader", action="store_false", dest="create_attention_mask_in_dataloader")
    _add(g, "--num-dataset-builder-threads", type=int, default=1)
    _add(g, "--s3-cache-path", type=str, default=None)
    g = p.add_argument_group("biencoder")       # reference _add_biencoder_args
    _add(g, "--ict-head-size", type=int, default=None)
    _add(g, "--biencoder-projection-dim", type=int, default=0)
    _add(g, "--biencoder-shared-query-context-model", action="store_true")
    _add(g, "--ict-load", type=str, default=None)
    _add(g, "--bert-load", type=str, default=None)
    _add(g, "--titles-data-path", type=str, default=None)
    _add(g, "--query-in-block-prob", type=float, default=0.1)
    _add(g, "--use-one-sent-docs", action="store_true")
    _add(g, "--evidence-data-path", type=str, default=None)
    _add(g, "--retriever-report-topk-accuracies", nargs="+", type=int, default=[])
    _add(g, "--retriever-score-scaling", action="store_true")
    _add(g, "--block-data-path", type=str, default=None)
    _add(g, "--embedding-path", type=str, default=None)
    _add(g, "--indexer-batch-size", type=int, default=128)
    _add(g, "--indexer-log-interval", type=int, default=1000)
    g = p.add_argument_group("reference CLI surface")   # flags of the reference parser that had no home above
    # -- deprecated / renamed (validate_args maps them) --
    _add(g, "--checkpoint-activations", action="store_true", help="deprecated: --recompute-granularity full --recompute-method uniform")
    _add(g, "--warmup", type=int, default=None, help="deprecated: use --lr-warmup-fraction")
    _add(g, "--inference-max-requests", type=int, default=8, dest="inference_max_requests")
    # -- model / heterogeneous / vision --
    _add(g, "--heterogeneous-layers-config-path", type=str, default=None)
    _add(g, "--heterogeneous-layers-config-encoded-json", type=str, default=None)
    _add(g, "--relative-attention-num-buckets", type=int, default=32)
    _add(g, "--relative-attention-max-distance", type=int, default=128)
    _add(g, "--mrope-section", type=int, nargs="+", default=None)
    _add(g, "--num-classes", type=int, default=1000)
    _add(g, "--img-h", type=int, default=224)
    _add(g, "--img-w", type=int, default=224)
    _add(g, "--num-channels", type=int, default=3)
    _add(g, "--patch-dim", type=int, default=16)
    _add(g, "--classes-fraction", type=float, default=1.0)
    _add(g, "--data-per-class-fraction", type=float, default=1.0)
    _add(g, "--no-data-sharding", action="store_false", dest="data_sharding")
    _add(g, "--head-lr-mult", type=float, default=1.0)
    _add(g, "--vision-pretraining", action="store_true")
    _add(g, "--vision-pretraining-type", type=str, default="classify", choices=["classify", "inpaint", "dino"])
    _add(g, "--vision-backbone-type", type=str, default="vit", choices=["vit", "mit", "swin"])
    _add(g, "--swin-backbone-type", type=str, default="tiny", choices=["tiny", "base", "h3"])
    _add(g, "--mask-type", type=str, default="random", choices=["random", "row"])
    _add(g, "--mask-factor", type=float, default=1.0)
    _add(g, "--iter-per-epoch", type=int, default=1250)
    for name, typ, dflt in (("local-img-size", int, 96), ("local-crops-number", int, 10), ("head-hidden-size", int, 2048),
                            ("bottleneck-size", int, 256), ("freeze-last-layer", float, 1), ("warmup-teacher-temp", float, 0.04),
                            ("teacher-temp", float, 0.07), ("warmup-teacher-temp-epochs", int, 30)):
        _add(g, f"--dino-{name}", type=typ, default=dflt)
    _add(g, "--dino-norm-last-layer", action="store_true")
    # -- retro --
    _add(g, "--retro-project-dir", type=str, default=None)
    _add(g, "--retro-add-retriever", action="store_true")
    _add(g, "--retro-cyclic-train-iters", type=int, default=None)
    _add(g, "--retro-encoder-layers", type=int, default=2)
    _add(g, "--retro-encoder-hidden-dropout", type=float, default=0.1)
    _add(g, "--retro-encoder-attention-dropout", type=float, default=0.1)
    # -- optimizer precision / offload --
    _add(g, "--use-precision-aware-optimizer", action="store_true")
    _add(g, "--main-grads-dtype", default="fp32", choices=["fp32", "bf16"])
    _add(g, "--main-params-dtype", default="fp32", choices=["fp32", "fp16"])
    _add(g, "--exp-avg-dtype", default="fp32", choices=["fp32", "fp16", "bf16", "fp8"])
    _add(g, "--exp-avg-sq-dtype", default="fp32", choices=["fp32", "fp16", "bf16", "fp8"])
    _add(g, "--grad-reduce-in-bf16", action="store_true")
    _add(g, "--optimizer-cpu-offload", action="store_true")
    _add(g, "--optimizer-offload-fraction", type=float, default=1.0)
    _add(g, "--use-torch-optimizer-for-cpu-offload", action="store_true")
    _add(g, "--overlap-cpu-optimizer-d2h-h2d", action="store_true")
    _add(g, "--no-pin-cpu-grads", action="store_false", dest="pin_cpu_grads")
    _add(g, "--no-pin-cpu-params", action="store_false", dest="pin_cpu_params")
    # -- distributed / runtime knobs (several are NCCL- or TE-specific in the reference; accepted, recorded in args) --
    _add(g, "--pipeline-model-parallel-comm-backend", type=str, default=None, choices=["nccl", "ucc"])
    _add(g, "--disable-gloo-process-groups", action="store_false", dest="enable_gloo_process_groups")
    _add(g, "--ddp-pad-buckets-for-high-nccl-busbw", action="store_true")
    _add(g, "--suggested-communication-unit-size", type=int, default=None)
    _add(g, "--no-gradient-reduce-div-fusion", action="store_false", dest="gradient_reduce_div_fusion")
    _add(g, "--no-clone-scatter-output-in-embedding", action="store_false", dest="clone_scatter_output_in_embedding")
    _add(g, "--disable-bf16-reduced-precision-matmul", action="store_true")
    _add(g, "--cross-entropy-fusion-impl", type=str, default="native", choices=["native", "te"])
    _add(g, "--cuda-graph-scope", type=str, default="full", choices=["full", "attn", "full_iteration"])
    _add(g, "--external-cuda-graph", action="store_true")
    _add(g, "--te-rng-tracker", action="store_true")
    _add(g, "--inference-rng-tracker", action="store_true")
    _add(g, "--init-model-with-meta-device", action="store_true")
    _add(g, "--keep-fp8-transpose-cache-when-using-custom-fsdp", action="store_true")
    _add(g, "--train-sync-interval", type=int, default=None)
    _add(g, "--run-workload-inspector-server", action="store_true")
    # -- checkpointing extras --
    _add(g, "--ckpt-convert-update-legacy-dist-opt-format", action="store_true")
    _add(g, "--use-mp-args-from-checkpoint-args", action="store_true")
    _add(g, "--no-use-tokenizer-model-from-checkpoint-args", action="store_false", dest="use_tokenizer_model_from_checkpoint_args")
    _add(g, "--use-persistent-ckpt-worker", action="store_true")
    _add(g, "--non-persistent-local-ckpt-dir", type=str, default=None)
    _add(g, "--non-persistent-local-ckpt-algo", type=str, default="fully_parallel", choices=["fully_parallel", "atomic"])
    _add(g, "--replication-jump", type=int, default=None)
    _add(g, "--replication-factor", type=int, default=2)
    _add(g, "--result-rejected-tracker-filename", type=str, default=None)
    # -- logging --
    _add(g, "--config-logger-dir", type=str, default="")
    _add(g, "--no-one-logger", action="store_false", dest="enable_one_logger")
    _add(g, "--one-logger-project", type=str, default="megatron-lm")
    _add(g, "--one-logger-run-name", type=str, default=None)
    _add(g, "--one-logger-async", action="store_true")
    _add(g, "--app-tag-run-name", type=str, default=None)
    _add(g, "--app-tag-run-version", type=str, default="0.0.0")
    g = p.add_argument_group("tokenizer")
    _add(g, "--vocab-size", type=int, default=None)
    _add(g, "--padded-vocab-size", type=int, default=None)
    _add(g, "--vocab-file", type=str, default=None)
    _add(g, "--merge-file", type=str, default=None)
    _add(g, "--vocab-extra-ids", type=int, default=0)
    _add(g, "--tokenizer-type", type=str, default=None)
    _add(g, "--tokenizer-model", type=str, default=None)
    _add(g, "--tiktoken-pattern", type=str, default=None)
    _add(g, "--tiktoken-num-special-tokens", type=int, default=1000)
    _add(g, "--tiktoken-special-tokens", type=str, nargs="+", default=None)
    _add(g, "--tokenizer-prompt-format", type=str, default=None)
    _add(g, "--special-tokens", type=str, nargs="*", default=None)
    _add(g, "--image-tag-type", type=str, default=None, choices=["nvlm", "internvl", ""])
    g = p.add_argument_group("logging")
    _add(g, "--log-params-norm", action="store_true")
    _add(g, "--log-num-zeros-in-grad", action="store_true")
    _add(g, "--log-throughput", action="store_true")
    _add(g, "--log-progress", action="store_true")
    _add(g, "--timing-log-level", type=int, default=0, choices=range(0, 3))
    _add(g, "--no-barrier-with-level-1-timing", action="store_false", dest="barrier_with_L1_time")
    _add(g, "--timing-log-option", type=str, default="minmax", choices=["max", "minmax", "all"])
    _add(g, "--tensorboard-log-interval", type=int, default=1)
    _add(g, "--tensorboard-queue-size", type=int, default=1000)
    _add(g, "--log-timers-to-tensorboard", action="store_true")
    _add(g, "--no-log-loss-scale-to-tensorboard", action="store_false", dest="log_loss_scale_to_tensorboard")
    _add(g, "--log-validation-ppl-to-tensorboard", action="store_true")
    _add(g, "--log-memory-to-tensorboard", action="store_true")
    _add(g, "--log-world-size-to-tensorboard", action="store_true")
    _add(g, "--wandb-project", type=str, default="")
    _add(g, "--wandb-exp-name", type=str, default="")
    _add(g, "--wandb-save-dir", type=str, default="")
    _add(g, "--logging-level", type=int, default=None)
    _add(g, "--log-straggler", action="store_true")
    _add(g, "--disable-straggler-on-startup", action="store_true")
    _add(g, "--straggler-ctrlr-port", type=int, default=65535)
    _add(g, "--straggler-minmax-count", type=int, default=1)
    g = p.add_argument_group("rerun / fault injection")
    _add(g, "--error-injection-rate", type=int, default=0)
    _add(g, "--error-injection-type", type=str, default="transient_error",
         choices=["correct_result", "transient_error", "persistent_error"])
    _add(g, "--rerun-mode", type=str, default="disabled", choices=["disabled", "validate_results", "report_stats"])
    _add(g, "--enable-ft-package", action="store_true")
    _add(g, "--calc-ft-timeouts", action="store_true")
    _add(g, "--simulated-fault", type=str, default=None,
         help="rank_hung|rank_killed;rank_to_fail;base_delay (MegaScan/FT test hook)")
    g = p.add_argument_group("inference")
    _add(g, "--inference-batch-times-seqlen-threshold", type=int, default=-1)
    _add(g, "--max-tokens-to-oom", type=int, default=12000)
    _add(g, "--output-bert-embeddings", action="store_true")
    _add(g, "--bert-embedder-type", default="megatron", choices=["megatron", "huggingface"])
    _add(g, "--flash-decode", action="store_true")
    _add(g, "--inference-max-seq-length", type=int, default=2560, dest="inference_max_seq_length")
    _add(g, "--inference-dynamic-batching", action="store_true")
    _add(g, "--inference-dynamic-batching-buffer-size-gb", type=float, default=40.0)
    _add(g, "--inference-dynamic-batching-buffer-guaranteed-fraction", type=float, default=0.2)
    _add(g, "--inference-dynamic-batching-buffer-overflow-factor", type=float, default=None)
    _add(g, "--inference-dynamic-batching-max-requests-override", type=int, default=None)
    _add(g, "--inference-dynamic-batching-max-tokens-override", type=int, default=None)
    _add(g, "--inference-dynamic-batching-chunk-size", type=int, default=256,
         help="tokens per KV block of the paged cache")
    _add(g, "--inference-max-batch-size", type=int, default=8, dest="inference_max_requests")
    g = p.add_argument_group("post-training")     # reference megatron/post_training/arguments.py (add_modelopt_args)
    _add(g, "--export-kd-cfg", type=str, default=None, help="distillation yaml: logit_layers, intermediate_layer_pairs, "
         "skip_lm_loss, kd_loss_scale")
    _add(g, "--export-kd-teacher-load", type=str, default=None, help="checkpoint of the (frozen) distillation teacher")
    _add(g, "--export-kd-finalize", action="store_true", help="load a distillation run and save the plain student")
    _add(g, "--export-model-type", type=str, default="GPTModel", choices=["GPTModel", "MambaModel"])
    _add(g, "--export-legacy-megatron", action="store_true")
    _add(g, "--export-te-mcore-model", action="store_true")
    _add(g, "--export-force-local-attention", action="store_true")
    _add(g, "--export-kv-cache-quant", action="store_true")
    _add(g, "--export-real-quant-cfg", type=str, default="None", choices=["fp8_real_quant", "fp8_blockwise_real_quant", "None"])
    _add(g, "--export-num-medusa-heads", type=int, default=0)
    _add(g, "--export-num-eagle-layers", type=int, default=0)
    _add(g, "--finetune-hf-dataset", type=str, default=None)
    g = p.add_argument_group("MegaScan")
    _add(g, "--trace", action="store_true")
    _add(g, "--trace-dir", type=str, default="trace")
    _add(g, "--trace-interval", type=int, default=5)
    _add(g, "--continuous-trace-iterations", type=int, default=2)
    _add(g, "--trace-granularity", type=str, default="full", choices=["base", "full"])
    _add(g, "--slow-rank-inject", type=str, default=None,
         help="software slow-rank injection '<global_rank>:<spin_cycles_per_layer>' (no root needed)")
    g = p.add_argument_group("MegaFBD")
    _add(g, "--forward-backward-disaggregating", action="store_true")
    _add(g, "--ignore-forward-tensor-parallel", action="store_true")
    g = p.add_argument_group("MegaDPP")
    _add(g, "--use-dpp", action="store_true")
    _add(g, "--dpp-policy", type=str, default="auto", choices=["auto", "bfc", "dfc", "static"])
    _add(g, "--node-ips", type=str, default="")
    _add(g, "--workload", type=int, default=2097152)
    _add(g, "--num-gpus", type=int, default=0)
    _add(g, "--multi-node", action="store_true")
    g = p.add_argument_group("MegaScope")
    _add(g, "--training-ws-port", type=int, default=None)
    _add(g, "--enable-ws-server", action="store_true")
    _add(g, "--ws-port", type=int, default=5000)
    _add(g, "--port", type=int, default=5000)
    _add(g, "--yaml-cfg", type=str, default=None)
    _add(g, "--spec", type=str, default=None, nargs="*")
    if extra_args_provider is not None:
        p = extra_args_provider(p)
    return p


def add_megatron_arguments(parser: argparse.ArgumentParser) -> argparse.ArgumentParser:
    """Add every framework argument to an existing parser (reference ``add_megatron_arguments``: embedding the training CLI in
    another program's parser)."""
    for action in build_parser()._actions:
        if action.dest == "help" or any(o in parser._option_string_actions for o in action.option_strings):
            continue
        parser._add_action(action)
    return parser


def moe_freq_type(x):
    """argparse type of ``--moe-layer-freq``: an int N (every N-th layer is MoE) or a python list expression like
    ``([1]*3+[0]*1)*3``; returned as int or list."""
    return _parse_moe_layer_freq(x) if isinstance(x, str) else x


def add_modelopt_args(parser):
    """The post-training (ModelOpt) flag group is part of the main parser here; kept for ``extra_args_provider=add_modelopt_args``."""
    return parser


def parse_args(extra_args_provider=None, ignore_unknown_args=False, args_list=None):
    parser = build_parser(extra_args_provider)
    if ignore_unknown_args:
        args, _ = parser.parse_known_args(args_list)
    else:
        args = parser.parse_args(args_list)
    args.rank = int(os.getenv("RANK", "0"))
    args.world_size = int(os.getenv("WORLD_SIZE", "1"))
    if args.yaml_cfg:
        import yaml

        with open(args.yaml_cfg) as f:
            for k, v in (yaml.safe_load(f) or {}).items():
                if isinstance(v, dict):
                    for kk, vv in v.items():
                        setattr(args, kk.replace("-", "_"), vv)
                else:
                    setattr(args, k.replace("-", "_"), v)
    return args


def _vocab_size_with_padding(orig: int, args) -> int:
    mult = args.make_vocab_size_divisible_by * args.tensor_model_parallel_size
    return (orig + mult - 1) // mult * mult


_UNSUPPORTED = (
    # (predicate on args, message) -- flags that are parsed for CLI compatibility but would silently change nothing
    (lambda a: any(getattr(a, k) != "fp32" for k in ("main_grads_dtype", "main_params_dtype")),
     "--main-grads-dtype / --main-params-dtype other than fp32 are not implemented (gradient accumulation and master "
     "weights stay fp32; the Adam moments may be fp16 / bf16 with --use-precision-aware-optimizer)"),
    (lambda a: "fp8" in (a.exp_avg_dtype, a.exp_avg_sq_dtype), "fp8 Adam moments are not implemented (use fp16 / bf16)"),
    (lambda a: not a.use_precision_aware_optimizer and (a.exp_avg_dtype != "fp32" or a.exp_avg_sq_dtype != "fp32"),
     "--exp-avg-dtype / --exp-avg-sq-dtype need --use-precision-aware-optimizer"),
    (lambda a: a.init_model_with_meta_device, "--init-model-with-meta-device is not implemented"),
    (lambda a: a.external_cuda_graph, "--external-cuda-graph is not implemented; use --enable-cuda-graph"),
    (lambda a: a.pipeline_model_parallel_comm_backend == "ucc", "the UCC pipeline backend is not available; p2p runs on NCCL/gloo channels"),
    (lambda a: a.non_persistent_ckpt_type == "local" and not a.non_persistent_local_ckpt_dir,
     "--non-persistent-ckpt-type local needs --non-persistent-local-ckpt-dir"),
    (lambda a: a.replication and (a.non_persistent_ckpt_type != "local" or not a.replication_jump),
     "--replication requires --non-persistent-ckpt-type local and --replication-jump"),
    (lambda a: a.non_persistent_ckpt_type == "in_memory", "in-memory non-persistent checkpoints are not implemented (use local / global)"),
    (lambda a: a.encoder_tensor_model_parallel_size not in (0, a.tensor_model_parallel_size),
     "a different tensor-parallel size for the encoder (--encoder-tensor-model-parallel-size) is not implemented"),
    (lambda a: a.optimizer_cpu_offload and not (0.0 < a.optimizer_offload_fraction <= 1.0),
     "--optimizer-offload-fraction must lie in (0, 1]"),
)


def validate_args(args, defaults=None):
    for bad, msg in _UNSUPPORTED:
        if bad(args):
            raise NotImplementedError(msg)
    if getattr(args, "ignore_forward_tensor_parallel", False) and not getattr(args, "forward_backward_disaggregating", False):
        raise ValueError("--ignore-forward-tensor-parallel needs --forward-backward-disaggregating")
    for k, v in (defaults or {}).items():
        if getattr(args, k, None) is None:
            setattr(args, k, v)
    if getattr(args, "checkpoint_activations", False):     # deprecated spelling of full uniform recompute
        args.recompute_granularity = args.recompute_granularity or "full"
        args.recompute_method = args.recompute_method or "uniform"
        args.recompute_num_layers = args.recompute_num_layers or 1
    if getattr(args, "warmup", None) is not None:
        raise ValueError("--warmup is no longer valid, use --lr-warmup-fraction instead")
    if args.model_parallel_size is not None:
        args.tensor_model_parallel_size = args.model_parallel_size
    if args.batch_size is not None and args.micro_batch_size is None:
        args.micro_batch_size = args.batch_size
    mp = args.tensor_model_parallel_size * args.pipeline_model_parallel_size * args.context_parallel_size
    collapse = bool(getattr(args, "ignore_forward_tensor_parallel", False) and args.forward_backward_disaggregating)
    if collapse:
        # every (stage, replica) is a block of ONE un-sharded forward rank + tp backward ranks
        blk = (args.tensor_model_parallel_size + 1) * args.pipeline_model_parallel_size
        if args.world_size % blk != 0:
            raise ValueError(f"world size ({args.world_size}) must be a multiple of (tp + 1) * pp = {blk} with "
                             "--ignore-forward-tensor-parallel")
        if args.context_parallel_size > 1 or args.expert_model_parallel_size > 1:
            raise ValueError("MegaFBD does not support context / expert parallelism")
        args.data_parallel_size = args.effective_data_parallel_size = args.world_size // blk
    elif args.world_size % mp != 0:
        raise ValueError(f"world size ({args.world_size}) is not divisible by tp*pp*cp ({mp})")
    else:
        args.data_parallel_size = args.world_size // mp
    if collapse:
        pass
    elif args.forward_backward_disaggregating:
        if args.data_parallel_size % 2 != 0:
            raise ValueError("--forward-backward-disaggregating needs an even data-parallel size")
        if args.context_parallel_size > 1 or args.expert_model_parallel_size > 1:
            raise ValueError("MegaFBD does not support context / expert parallelism")
        args.effective_data_parallel_size = args.data_parallel_size // 2
    else:
        args.effective_data_parallel_size = args.data_parallel_size
    if args.micro_batch_size is None:
        raise ValueError("--micro-batch-size is required")
    if args.global_batch_size is None:
        args.global_batch_size = args.micro_batch_size * args.effective_data_parallel_size
    if args.global_batch_size % (args.micro_batch_size * args.effective_data_parallel_size) != 0:
        raise ValueError("global batch size must be divisible by micro batch size * data parallel size")
    # virtual pipeline
    args.virtual_pipeline_model_parallel_size = None
    if args.num_layers_per_virtual_pipeline_stage is not None:
        if args.pipeline_model_parallel_size < 2:
            raise ValueError("interleaved schedule needs pipeline-model-parallel size >= 2")
        per_stage = args.num_layers // args.pipeline_model_parallel_size
        if per_stage % args.num_layers_per_virtual_pipeline_stage != 0:
            raise ValueError("layers per pipeline stage not divisible by layers per virtual stage")
        args.virtual_pipeline_model_parallel_size = per_stage // args.num_layers_per_virtual_pipeline_stage
    elif args.num_virtual_stages_per_pipeline_rank is not None:
        args.virtual_pipeline_model_parallel_size = args.num_virtual_stages_per_pipeline_rank
    if args.virtual_pipeline_model_parallel_size == 1:
        args.virtual_pipeline_model_parallel_size = None
    # dtypes
    args.params_dtype = torch.float32
    if args.fp16:
        if args.bf16:
            raise ValueError("--fp16 and --bf16 are exclusive")
        args.params_dtype = torch.half
    if args.bf16:
        args.params_dtype = torch.bfloat16
        # --grad-reduce-in-bf16 (reference arguments.py: bf16 reduction instead of fp32).  With the fused reduce-scatter kernel
        # the main grads stay fp32 and only the NVLink wire format is bf16; with NCCL the grad buffers themselves are bf16.
        args.accumulate_allreduce_grads_in_fp32 = not (args.grad_reduce_in_bf16 and not args.fused_grad_reduce_scatter)
    if args.dataloader_type is None:
        args.dataloader_type = "single"
    args.consumed_train_samples = 0
    args.skipped_train_samples = 0
    args.consumed_valid_samples = 0
    if args.train_iters is not None and args.train_samples is not None:
        raise ValueError("only one of --train-iters / --train-samples")
    if args.train_samples is not None and args.train_iters is None:
        args.train_iters = args.train_samples // args.global_batch_size
    if args.lr_decay_iters is None and args.train_iters is not None:
        args.lr_decay_iters = args.train_iters
    if args.lr_warmup_fraction is not None and args.lr_warmup_iters != 0:
        raise ValueError("only one of --lr-warmup-fraction / --lr-warmup-iters")
    for req in ("num_layers", "hidden_size", "num_attention_heads"):
        if getattr(args, req) is None:
            raise ValueError(f"--{req.replace('_', '-')} is required")
    if args.max_position_embeddings is None:
        args.max_position_embeddings = args.seq_length
    if args.seq_length is None and args.encoder_seq_length is not None:
        args.seq_length = args.encoder_seq_length
    if args.encoder_seq_length is None:
        args.encoder_seq_length = args.seq_length
    if args.seq_length is None:
        raise ValueError("--seq-length is required")
    if args.position_embedding_type == "learned_absolute" and args.max_position_embeddings < args.seq_length:
        raise ValueError("max_position_embeddings must be >= seq_length")
    if args.use_rotary_position_embeddings:
        args.position_embedding_type = "rope"
    if args.ffn_hidden_size is None:
        if args.swiglu:
            args.ffn_hidden_size = int((4 * args.hidden_size * 2 / 3) / 64) * 64
        else:
            args.ffn_hidden_size = 4 * args.hidden_size
    if args.kv_channels is None:
        args.kv_channels = args.hidden_size // args.num_attention_heads
    if not args.group_query_attention:
        args.num_query_groups = args.num_attention_heads
    if args.layernorm_epsilon is not None:
        args.norm_epsilon = args.layernorm_epsilon
    if args.start_weight_decay is None:
        args.start_weight_decay = args.weight_decay
        args.end_weight_decay = args.weight_decay
    if args.sequence_parallel and args.tensor_model_parallel_size == 1:
        args.sequence_parallel = False
    if args.context_parallel_size > 1 and args.seq_length % (2 * args.context_parallel_size) != 0:
        raise ValueError("seq-length must be divisible by 2*context-parallel-size")
    if args.tp_comm_overlap and not args.sequence_parallel:
        raise ValueError("--tp-comm-overlap requires --sequence-parallel")
    if args.recompute_activations:
        args.recompute_granularity = "selective"
    if args.recompute_granularity == "selective" and args.recompute_modules is None:
        args.recompute_modules = ["core_attn"]
    if args.use_distributed_optimizer is False and args.overlap_param_gather:
        raise ValueError("--overlap-param-gather needs --use-distributed-optimizer")
    if args.num_experts is not None and args.num_experts % args.expert_model_parallel_size != 0:
        raise ValueError("num-experts must be divisible by expert-model-parallel-size")
    if args.expert_model_parallel_size > 1 and args.num_experts is None:
        raise ValueError("--num-experts is required for expert parallelism")
    if isinstance(args.moe_layer_freq, str):
        args.moe_layer_freq = _parse_moe_layer_freq(args.moe_layer_freq)
    if args.use_dpp and args.virtual_pipeline_model_parallel_size is None and args.pipeline_model_parallel_size > 1:
        args.use_dpp_static_ok = True
    if args.ckpt_format == "torch" and args.async_save:
        raise ValueError("--async-save needs a distributed checkpoint format")
    if args.forward_backward_disaggregating:
        args.ckpt_format = "torch"  # per-rank .pt, as in the reference (training.py:833-834)
    if args.use_cpu_initialization is None:
        args.use_cpu_initialization = False
    if args.save_interval is None and args.save:
        args.save_interval = args.train_iters
    args.padded_vocab_size = getattr(args, "padded_vocab_size", None)
    return args


def _parse_moe_layer_freq(s: str):
    s = s.strip()
    if s.isdigit():
        return int(s)
    out = eval(s, {"__builtins__": {}}, {})  # noqa: S307 - list-literal arithmetic like ([1]*3+[0])*2
    if isinstance(out, (list, tuple)):
        return [int(x) for x in out]
    return int(out)


def core_transformer_config_from_args(args, config_class=None) -> TransformerConfig:
    config_class = config_class or (MLATransformerConfig if args.multi_latent_attention else TransformerConfig)
    kw = {}
    for f in dataclasses.fields(config_class):
        if hasattr(args, f.name):
            kw[f.name] = getattr(args, f.name)
    from ..parallel import state as _ps

    if _ps.fbd_forward_tp_collapsed() and _ps.is_forward_stage():
        # un-sharded forward rank: true TP=1 execution; remember what its backward twins run so that the stage inputs it
        # hands over have their layout (sequence-parallel slices)
        import types

        kw["fbd_backward_config"] = types.SimpleNamespace(sequence_parallel=bool(kw.get("sequence_parallel")),
                                                          tensor_model_parallel_size=kw.get("tensor_model_parallel_size", 1))
        kw["tensor_model_parallel_size"] = 1
        kw["sequence_parallel"] = False
    kw["persist_layer_norm"] = not args.no_persist_layer_norm
    kw["layernorm_zero_centered_gamma"] = args.apply_layernorm_1p
    kw["layernorm_epsilon"] = args.norm_epsilon
    kw["deallocate_pipeline_outputs"] = True
    kw["pipeline_dtype"] = args.params_dtype
    kw["batch_p2p_comm"] = not args.overlap_p2p_comm
    kw["num_moe_experts"] = args.num_experts
    kw["rotary_interleaved"] = args.rotary_interleaved
    kw["num_layers_in_first_pipeline_stage"] = args.decoder_first_pipeline_num_layers
    kw["num_layers_in_last_pipeline_stage"] = args.decoder_last_pipeline_num_layers
    kw["tp_comm_overlap"] = bool(args.tp_comm_overlap)
    if getattr(args, "microbatch_group_size_per_virtual_pipeline_stage", None):
        kw["microbatch_group_size_per_vp_stage"] = args.microbatch_group_size_per_virtual_pipeline_stage
    kw["bias_activation_fusion"] = False
    if args.swiglu:
        kw["activation_func"] = F.silu
        kw["gated_linear_unit"] = True
    elif args.squared_relu:
        from ..transformer.mlp import squared_relu

        kw["activation_func"] = squared_relu
    elif args.onnx_safe and not args.openai_gelu:
        from ..transformer.mlp import erf_gelu

        kw["activation_func"] = erf_gelu        # --openai-gelu (tanh form) is what the default fused bias-GELU computes
    if args.init_method_xavier_uniform:
        kw["init_method"] = torch.nn.init.xavier_uniform_
    if args.group_query_attention:
        kw["num_query_groups"] = args.num_query_groups
    else:
        kw["num_query_groups"] = None
    kw["config_logger_dir"] = getattr(args, "config_logger_dir", "") or ""
    if len(args.cp_comm_type) == 1:
        kw["cp_comm_type"] = args.cp_comm_type[0]
    kw["kernel_backend"] = args.kernel_backend
    return config_class(**kw)
