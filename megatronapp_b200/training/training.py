"""``pretrain()`` and the training loop.

Parity: reference ``megatron/training/training.py`` -- ``pretrain`` :894 (takes
``train_valid_test_datasets_provider, model_provider, model_type, forward_step_func``),
``get_model`` :1043 (one module per (virtual) pipeline chunk, ``Float16Module``, DDP wrap,
``broadcast_params``), ``setup_model_and_optimizer`` :1259, ``train_step`` :1367,
``train`` :1967, ``training_log`` :1488, ``num_floating_point_operations`` :142-395,
``evaluate`` :2371, checkpoint/exit policy :1803-1965; MegaScan wiring :2167-2171,2249,2339;
MegaScope training server :1975-2023; MegaDPP / MegaFBD schedule selection.
"""
from __future__ import annotations

import gc
import math
import os
import sys
import time
from typing import Callable, List, Optional

import torch
import torch.distributed as dist

from ..optim import OptimizerConfig, OptimizerParamScheduler, get_megatron_optimizer
from ..parallel import p2p, schedules
from ..parallel import state as ps
from ..parallel.ddp import DistributedDataParallel, DistributedDataParallelConfig, finalize_model_grads
from ..scan.tracer import get_tracer
from ..transformer.module import Float16Module
from ..utils.microbatches import get_current_global_batch_size, get_num_microbatches, update_num_microbatches
from ..utils.straggler import get_straggler_detector
from . import checkpointing, ft_integration
from .rerun_state_machine import (RerunDataIterator, RerunErrorInjector, get_rerun_state_machine,
                                  initialize_rerun_state_machine)
from .arguments import core_transformer_config_from_args
from .global_vars import get_args, get_tensorboard_writer, get_timers, get_signal_handler
from .initialize import initialize_megatron
from .utils import (calc_params_l2_norm, is_last_rank, print_rank_0, print_rank_last, report_memory,
                    reduce_max_stat_across_model_parallel_group, logical_and_across_model_parallel_group,
                    unwrap_model)

_TRAIN_START_TIME = time.time()


# ---------------------------------------------------------------------------
# analytic FLOP model (dense / GQA / SwiGLU / MoE)
# ---------------------------------------------------------------------------

def num_floating_point_operations(args, batch_size: int) -> float:
    """fwd+bwd FLOPs of one global batch (3x forward GEMM FLOPs; reference :142-395)."""
    h, s, L, V = args.hidden_size, args.seq_length, args.num_layers, args.padded_vocab_size
    kv = args.kv_channels * args.num_query_groups
    q = args.kv_channels * args.num_attention_heads
    gated = 2 if args.swiglu else 1
    n_exp_layers = 0
    if args.num_experts:
        freq = args.moe_layer_freq
        n_exp_layers = L // freq if isinstance(freq, int) else sum(freq)
    dense_layers = L - n_exp_layers
    ffn = args.ffn_hidden_size
    moe_ffn = args.moe_ffn_hidden_size or ffn
    shared = args.moe_shared_expert_intermediate_size or 0
    attn = 2 * s * h * (q + 2 * kv) + 2 * s * q * h          # qkv + proj GEMMs per sample
    core = 2 * 2 * s * s * q / (2 if True else 1)            # QK^T and PV with causal half
    mlp_dense = 2 * s * h * ffn * (gated + 1)
    mlp_moe = 2 * s * h * (moe_ffn * args.moe_router_topk + shared) * (gated + 1) if args.num_experts else 0
    logits = 2 * s * h * V
    per_sample = L * (attn + core) + dense_layers * mlp_dense + n_exp_layers * mlp_moe + logits
    return 3.0 * batch_size * per_sample


# ---------------------------------------------------------------------------
# model construction
# ---------------------------------------------------------------------------

def get_model(model_provider_func: Callable, model_type=None, wrap_with_ddp: bool = True):
    args = get_args()
    pp = ps.get_pipeline_model_parallel_world_size()
    vpp = args.virtual_pipeline_model_parallel_size
    model = []
    if pp > 1 and vpp is not None:
        for i in range(vpp):
            ps.set_virtual_pipeline_model_parallel_rank(i)
            m = model_provider_func(pre_process=ps.is_pipeline_first_stage(), post_process=ps.is_pipeline_last_stage())
            m.model_type = model_type
            model.append(m)
        ps.set_virtual_pipeline_model_parallel_rank(0)
    elif pp > 1 and str(model_type) == "encoder_and_decoder":
        # encoder stages [0, split), decoder stages [split, pp): embeddings on the first stage of each half, heads /
        # final norms on the last stage of each half (reference training.py get_model, encoder_and_decoder branch)
        if vpp is not None:
            raise ValueError("interleaved schedule is not supported for encoder-decoder models")
        rank, split = ps.get_pipeline_model_parallel_rank(), ps.get_pipeline_model_parallel_split_rank()
        if split is None:
            raise ValueError("encoder-decoder model with pp > 1 needs --pipeline-model-parallel-split-rank")
        m = model_provider_func(pre_process=rank in (0, split), post_process=rank in (split - 1, pp - 1),
                                add_encoder=ps.is_inside_encoder(rank), add_decoder=ps.is_inside_decoder(rank))
        m.model_type = model_type
        model = [m]
    else:
        m = model_provider_func(pre_process=ps.is_pipeline_first_stage(), post_process=ps.is_pipeline_last_stage())
        m.model_type = model_type
        model = [m]
    for m in model:
        for p in m.parameters():
            if not hasattr(p, "tensor_model_parallel"):
                p.tensor_model_parallel = False
    if ps.get_data_parallel_rank() == 0:
        n = sum(sum(p.nelement() for p in m.parameters()) for m in model)
        print(f" > number of parameters on (tensor, pipeline) model parallel rank "
              f"({ps.get_tensor_model_parallel_rank()}, {ps.get_pipeline_model_parallel_rank()}): {n}", flush=True)
    if torch.cuda.is_available() and not args.use_cpu_initialization:
        for m in model:
            m.cuda(torch.cuda.current_device())
    if args.fp16 or args.bf16:
        config = unwrap_model(model[0]).config
        model = [Float16Module(config, m) for m in model]
    if wrap_with_ddp:
        config = unwrap_model(model[0]).config
        ddp_config = DistributedDataParallelConfig(
            grad_reduce_in_fp32=args.accumulate_allreduce_grads_in_fp32,
            overlap_grad_reduce=args.overlap_grad_reduce,
            overlap_param_gather=args.overlap_param_gather,
            use_distributed_optimizer=args.use_distributed_optimizer,
            num_distributed_optimizer_instances=args.num_distributed_optimizer_instances,
            check_for_nan_in_grad=False,
            check_for_large_grads=args.check_for_large_grads,
            bucket_size=args.ddp_bucket_size,
            average_in_collective=args.ddp_average_in_collective,
            fused_grad_reduce_scatter=args.fused_grad_reduce_scatter,
            grad_wire_dtype="bf16" if args.grad_reduce_in_bf16 else "fp32")
        if args.use_custom_fsdp or args.use_torch_fsdp2:
            from ..parallel.fsdp import FullyShardedDataParallel, TorchFullyShardedDataParallel

            cls = FullyShardedDataParallel if args.use_custom_fsdp else TorchFullyShardedDataParallel
            model = [cls(config, ddp_config, m) for m in model]
        else:
            model = [DistributedDataParallel(config, ddp_config, m, disable_bucketing=(i > 0))
                     for i, m in enumerate(model)]
        if args.data_parallel_random_init:
            for m in model:
                m.broadcast_params()
    return model


def get_optimizer_param_scheduler(optimizer):
    args = get_args()
    if args.lr is None:
        raise ValueError("--lr is required to train (no learning rate was given)")
    if args.train_iters:
        if args.lr_decay_iters is None:
            args.lr_decay_iters = args.train_iters
        lr_decay_steps = args.lr_decay_iters * args.global_batch_size
        wd_incr_steps = args.train_iters * args.global_batch_size
        wsd_decay_steps = args.lr_wsd_decay_iters * args.global_batch_size if args.lr_wsd_decay_iters else None
        if args.lr_warmup_fraction is not None:
            lr_warmup_steps = args.lr_warmup_fraction * lr_decay_steps
        else:
            lr_warmup_steps = args.lr_warmup_iters * args.global_batch_size
    elif args.train_samples:
        lr_decay_steps = args.lr_decay_samples or args.train_samples
        wd_incr_steps = args.train_samples
        wsd_decay_steps = args.lr_wsd_decay_samples
        lr_warmup_steps = (args.lr_warmup_fraction * lr_decay_steps if args.lr_warmup_fraction is not None
                           else args.lr_warmup_samples)
    else:
        raise ValueError("either --train-iters or --train-samples must be given")
    return OptimizerParamScheduler(
        optimizer, init_lr=args.lr_warmup_init, max_lr=args.lr, min_lr=args.min_lr,
        lr_warmup_steps=lr_warmup_steps, lr_decay_steps=lr_decay_steps, lr_decay_style=args.lr_decay_style,
        start_wd=args.start_weight_decay, end_wd=args.end_weight_decay, wd_incr_steps=wd_incr_steps,
        wd_incr_style=args.weight_decay_incr_style,
        use_checkpoint_opt_param_scheduler=args.use_checkpoint_opt_param_scheduler,
        override_opt_param_scheduler=args.override_opt_param_scheduler,
        wsd_decay_steps=wsd_decay_steps, lr_wsd_decay_style=args.lr_wsd_decay_style)


def _upcycle_dense_checkpoint(args, model_provider_func, model_type, model, optimizer):
    """--moe-use-upcycling (reference training.py:1284-1311): build the dense twin of this MoE model, load the dense checkpoint
    from --load into it, seed every expert from the dense MLP, save the result as iteration 1 under --save and start from
    iteration 0.  Only for the first run of a conversion: a --save directory that already holds a checkpoint is an error."""
    from ..moe import upcycling_utils

    if dist.is_initialized():
        dist.barrier()
    if args.save and os.path.isfile(checkpointing.get_checkpoint_tracker_filename(args.save)):
        raise RuntimeError("the upcycling destination already holds a checkpoint: --moe-use-upcycling is only for the first "
                           "run that converts the dense model; remove the flag to resume")
    if not args.load:
        raise ValueError("--moe-use-upcycling needs --load <dense checkpoint>")
    keep = (args.num_experts, args.expert_model_parallel_size, args.finetune, args.no_load_rng)
    args.num_experts, args.expert_model_parallel_size, args.finetune, args.no_load_rng = None, 1, True, True
    try:
        dense = get_model(model_provider_func, model_type, wrap_with_ddp=False)
        _, flops = upcycling_utils.load_and_upcycle_model(
            checkpointing.load_checkpoint, unwrap_model(model), dense,
            load_kwargs={"model": dense, "optimizer": None, "opt_param_scheduler": None, "load_arg": "load"})
    finally:
        args.num_experts, args.expert_model_parallel_size, args.finetune, args.no_load_rng = keep
    del dense
    if optimizer is not None and hasattr(optimizer, "reload_model_params"):
        optimizer.reload_model_params()
    args.num_floating_point_operations_so_far = flops
    if args.save:
        checkpointing.save_checkpoint(1, model, None, None, flops)
        checkpointing.maybe_finalize_async_save(blocking=True)
        print_rank_0(f"upcycled checkpoint saved to {args.save}")
    args.iteration = 0


def setup_model_and_optimizer(model_provider_func, model_type=None, no_wd_decay_cond=None, scale_lr_cond=None,
                              lr_mult=1.0):
    args = get_args()
    model = get_model(model_provider_func, model_type)
    kw = {}
    import dataclasses

    for f in dataclasses.fields(OptimizerConfig):
        if hasattr(args, f.name):
            kw[f.name] = getattr(args, f.name)
    names = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}
    for k in ("main_grads_dtype", "main_params_dtype", "exp_avg_dtype", "exp_avg_sq_dtype"):
        if isinstance(kw.get(k), str):
            kw[k] = names[kw[k]]
    config = OptimizerConfig(**kw)
    config.timers = get_timers()
    optimizer = get_megatron_optimizer(config, model, no_wd_decay_cond, scale_lr_cond, lr_mult)
    scheduler = get_optimizer_param_scheduler(optimizer)
    if args.moe_use_upcycling:
        _upcycle_dense_checkpoint(args, model_provider_func, model_type, model, optimizer)
    elif args.load is not None or args.pretrained_checkpoint is not None:
        load_arg = "load" if (args.load and os.path.isfile(checkpointing.get_checkpoint_tracker_filename(args.load))) \
            else ("pretrained_checkpoint" if args.pretrained_checkpoint else "load")
        args.iteration, args.num_floating_point_operations_so_far = checkpointing.load_checkpoint(
            model, optimizer, scheduler, load_arg=load_arg)
    else:
        args.iteration, args.num_floating_point_operations_so_far = 0, 0
    core = unwrap_model(model)
    if args.iteration == 0 and len(core) == 1 and hasattr(core[0], "init_state_dict_from_bert") and getattr(args, "bert_load", None):
        # ICT / retriever pre-training starts from a pre-trained BERT (reference training.py:1333-1338)
        print_rank_0("initializing the bi-encoder from the pretrained BERT model")
        core[0].init_state_dict_from_bert()
        if hasattr(optimizer, "reload_model_params"):
            optimizer.reload_model_params()
    # TP fused-comm arena (fused AG->GEMM / GEMM->RS kernels)
    if (args.sequence_parallel and args.fused_tp_comm and torch.cuda.is_available()
            and ps.get_tensor_model_parallel_world_size() > 1 and args.kernel_backend != "torch"):
        from ..parallel.tp_layers import setup_tp_fused_comm

        max_out = max(args.ffn_hidden_size * (2 if args.swiglu else 1) // args.tensor_model_parallel_size,
                      args.hidden_size)
        setup_tp_fused_comm(args.seq_length * args.micro_batch_size, args.hidden_size, max_out)
    if args.fused_grad_reduce_scatter and torch.cuda.is_available() and args.use_distributed_optimizer:
        from ..comm.arena import get_arena

        dp_group = ps.get_data_parallel_group(with_context_parallel=True)
        if dist.get_world_size(dp_group) > 1:
            mx = max(buf.max_bucket_numel() for m in model for buf in m.buffers)
            get_arena("dp", dp_group, 2 * mx * 4 + (1 << 20))
    return model, optimizer, scheduler


# ---------------------------------------------------------------------------
# one training step
# ---------------------------------------------------------------------------

def train_step(forward_step_func, data_iterator, model, optimizer, opt_param_scheduler, config):
    args = get_args()
    timers = get_timers()
    tracer = get_tracer()
    for m in model:
        m.zero_grad_buffer()
    optimizer.zero_grad()
    fb = schedules.get_forward_backward_func()
    rsm = get_rerun_state_machine()
    first = True
    losses_reduced = []
    # the state machine may ask for the same step to be re-executed on identical data (fault
    # classification, reference training.py:1322 / rerun_state_machine.py:258)
    while rsm.should_run_forward_backward(data_iterator):
        if not first:
            for m in model:
                m.zero_grad_buffer()
            optimizer.zero_grad()
        first = False
        losses_reduced = fb(forward_step_func=forward_step_func, data_iterator=data_iterator, model=model,
                            num_microbatches=get_num_microbatches(), seq_length=args.seq_length,
                            micro_batch_size=args.micro_batch_size, decoder_seq_length=args.decoder_seq_length,
                            forward_only=False)
    should_ckpt, should_exit, exit_code = rsm.should_checkpoint_and_exit()
    if should_exit:
        return {}, True, should_ckpt, should_exit, exit_code, None, None
    if args.empty_unused_memory_level >= 1:
        torch.cuda.empty_cache()
    timers("optimizer", log_level=1).start(barrier=args.barrier_with_L1_time)
    if args.forward_backward_disaggregating and ps.is_forward_stage():
        update_successful, grad_norm, num_zeros_in_grad = True, None, None   # twin owns the optimizer
    else:
        with tracer.scope("optimizer"):
            update_successful, grad_norm, num_zeros_in_grad = optimizer.step()
    if args.forward_backward_disaggregating:
        from ..fbd.schedule import sync_forward_weights

        sync_forward_weights(model)
    timers("optimizer").stop()
    update_successful = logical_and_across_model_parallel_group(update_successful)
    grad_norm = reduce_max_stat_across_model_parallel_group(grad_norm)
    if update_successful and getattr(args, "vision_pretraining", False) and args.vision_pretraining_type == "dino":
        from .utils import unwrap_model

        unwrap_model(model[0]).update_momentum(args.curr_iteration)     # EMA teacher (reference training.py train_step)
    if update_successful:
        increment = get_num_microbatches() * args.micro_batch_size * args.effective_data_parallel_size
        opt_param_scheduler.step(increment=increment)
        skipped_iter = 0
    else:
        skipped_iter = 1
    if args.empty_unused_memory_level >= 2:
        torch.cuda.empty_cache()
    loss_reduced = {}
    if ps.is_pipeline_last_stage(ignore_virtual=True) and losses_reduced:
        for key in losses_reduced[0].keys():
            vals = [x[key] for x in losses_reduced]
            if torch.is_tensor(vals[0]) and vals[0].numel() == 2:
                v = torch.stack(vals).sum(dim=0)
                loss_reduced[key] = v[0] / v[1]
            elif isinstance(vals[0], (tuple, list)) and len(vals[0]) == 2:
                num = sum(v[0] for v in vals)
                den = sum(v[1] for v in vals)
                loss_reduced[key] = num / den
            else:
                loss_reduced[key] = sum(vals) / len(vals)
    return loss_reduced, skipped_iter, False, False, 0, grad_norm, num_zeros_in_grad


def training_log(loss_dict, total_loss_dict, learning_rate, iteration, loss_scale, report_memory_flag, skipped_iter,
                 grad_norm, params_norm, num_zeros_in_grad, elapsed_per_iter=None):
    args = get_args()
    timers = get_timers()
    writer = get_tensorboard_writer()
    adv, skip, nan = "advanced iterations", "skipped iterations", "nan iterations"
    total_loss_dict[adv] = total_loss_dict.get(adv, 0) + (0 if skipped_iter else 1)
    total_loss_dict[skip] = total_loss_dict.get(skip, 0) + skipped_iter
    got_nan = False
    for k, v in loss_dict.items():
        if not skipped_iter:
            fv = float(v)
            total_loss_dict[k] = total_loss_dict.get(k, 0.0) + fv
            got_nan = got_nan or not math.isfinite(fv)
    total_loss_dict[nan] = total_loss_dict.get(nan, 0) + int(got_nan)
    # auxiliary losses that never reach the loss function's report: MoE load balancing / z-loss (per layer, summed over
    # micro-batches) and the multi-token-prediction heads (reference training_log: track_moe_metrics, track_mtp_metrics)
    if args.num_experts is not None:
        from ..moe.moe_utils import track_moe_metrics

        track_moe_metrics(1.0 / get_num_microbatches(), iteration, writer, None, None if skipped_iter else total_loss_dict,
                          getattr(args, "moe_per_layer_logging", False))
    if getattr(args, "mtp_num_layers", None):
        from ..transformer.multi_token_prediction import MTPLossLoggingHelper

        MTPLossLoggingHelper.track_mtp_metrics(1.0 / get_num_microbatches(), iteration, writer,
                                               None if skipped_iter else total_loss_dict)
    batch_size = args.micro_batch_size * args.effective_data_parallel_size * get_num_microbatches()
    if writer and iteration % args.tensorboard_log_interval == 0:
        writer.add_scalar("learning-rate", learning_rate, iteration)
        writer.add_scalar("batch-size", batch_size, iteration)
        for k, v in loss_dict.items():
            writer.add_scalar(k, float(v), iteration)
            writer.add_scalar(k + " vs samples", float(v), args.consumed_train_samples)
        if args.log_loss_scale_to_tensorboard:
            writer.add_scalar("loss-scale", loss_scale, iteration)
        if grad_norm is not None:
            writer.add_scalar("grad-norm", grad_norm, iteration)
        if num_zeros_in_grad is not None:
            writer.add_scalar("num-zeros", num_zeros_in_grad, iteration)
        if params_norm is not None:
            writer.add_scalar("params-norm", params_norm, iteration)
        if args.log_memory_to_tensorboard and torch.cuda.is_available():
            ms = torch.cuda.memory_stats()
            writer.add_scalar("mem-reserved-bytes", ms["reserved_bytes.all.current"], iteration)
            writer.add_scalar("mem-allocated-bytes", ms["allocated_bytes.all.current"], iteration)
    if iteration % args.log_interval == 0:
        elapsed = timers("interval-time").elapsed(barrier=True)
        per_iter = elapsed / max(1, total_loss_dict[adv] + total_loss_dict[skip])
        throughput = num_floating_point_operations(args, batch_size) / (per_iter * 1e12 * args.world_size) if per_iter > 0 else 0.0
        tokens_per_sec = batch_size * args.seq_length / per_iter if per_iter > 0 else 0.0
        if writer:
            writer.add_scalar("iteration-time", per_iter, iteration)
            writer.add_scalar("throughput", throughput, iteration)
            if args.log_world_size_to_tensorboard:
                writer.add_scalar("world-size", args.world_size, iteration)
            if args.log_timers_to_tensorboard:
                timers.write([n for n in timers._timers if n != "interval-time"], writer, iteration,
                             normalizer=args.log_interval, reset=False)
        s = f" [{time.strftime('%Y-%m-%d %H:%M:%S')}]"
        s += f" iteration {iteration:8d}/{args.train_iters:8d} |"
        s += f" consumed samples: {args.consumed_train_samples:12d} |"
        s += f" elapsed time per iteration (ms): {per_iter * 1000.0:.1f} |"
        if args.log_throughput:
            s += f" throughput per GPU (TFLOP/s/GPU): {throughput:.1f} |"
            s += f" tokens/s: {tokens_per_sec:.0f} |"
        s += f" learning rate: {learning_rate:.6E} |"
        s += f" global batch size: {batch_size:5d} |"
        for k in list(total_loss_dict):
            if k in (adv, skip, nan):
                continue
            avg = total_loss_dict[k] / float(max(1, total_loss_dict[adv]))
            if avg > 0.0:
                s += f" {k}: {avg:.6E} |"
            total_loss_dict[k] = 0.0
        s += f" loss scale: {loss_scale:.1f} |"
        if grad_norm is not None:
            s += f" grad norm: {grad_norm:.3f} |"
        if num_zeros_in_grad is not None:
            s += f" num zeros: {num_zeros_in_grad} |"
        if params_norm is not None:
            s += f" params norm: {params_norm:.3f} |"
        s += f" number of skipped iterations: {total_loss_dict[skip]:3d} |"
        s += f" number of nan iterations: {total_loss_dict[nan]:3d} |"
        total_loss_dict[adv] = total_loss_dict[skip] = total_loss_dict[nan] = 0
        print_rank_last(s)
        if report_memory_flag:
            report_memory(f"(after {iteration} iterations)")
            report_memory_flag = False
        timers.log(None, normalizer=args.log_interval)
    return report_memory_flag


def evaluate(forward_step_func, data_iterator, model, config, verbose=False):
    args = get_args()
    for m in model:
        m.eval()
    total = {}
    fb = schedules.get_forward_backward_func()
    n = 0
    with torch.no_grad():
        for it in range(args.eval_iters):
            ft_integration.on_eval_step_start()
            out = fb(forward_step_func=forward_step_func, data_iterator=data_iterator, model=model,
                     num_microbatches=get_num_microbatches(), seq_length=args.seq_length,
                     micro_batch_size=args.micro_batch_size, forward_only=True)
            ft_integration.on_eval_step_end()
            if ps.is_pipeline_last_stage(ignore_virtual=True):
                for d in out:
                    for k, v in d.items():
                        if torch.is_tensor(v) and v.numel() == 2:
                            acc = total.get(k, torch.zeros(2, device=v.device))
                            total[k] = acc + v
                        elif isinstance(v, (tuple, list)) and len(v) == 2:
                            acc = total.get(k, torch.zeros(2))
                            total[k] = acc + torch.tensor([float(v[0]), float(v[1])])
                        else:
                            acc = total.get(k, torch.zeros(2))
                            total[k] = acc + torch.tensor([float(v), 1.0])
            args.consumed_valid_samples += get_current_global_batch_size()
            n += 1
    for m in model:
        m.train()
    return {k: float(v[0] / v[1]) for k, v in total.items()}, None, False


def evaluate_and_print_results(prefix, forward_step_func, data_iterator, model, iteration, config, verbose=False):
    res, _, _ = evaluate(forward_step_func, data_iterator, model, config, verbose)
    s = f" validation loss at {prefix} | "
    writer = get_tensorboard_writer()
    for k, v in res.items():
        s += f"{k} value: {v:.6E} | "
        ppl = math.exp(min(20, v))
        s += f"{k} PPL: {ppl:.6E} | "
        if writer:
            writer.add_scalar(f"{k} validation", v, iteration)
            if get_args().log_validation_ppl_to_tensorboard:
                writer.add_scalar(f"{k} validation ppl", ppl, iteration)
    print_rank_last("-" * (len(s) + 1))
    print_rank_last(s)
    print_rank_last("-" * (len(s) + 1))


def append_to_progress_log(line: str, barrier: bool = True):
    """``--log-progress``: one tab-separated line per event in ``<save>/progress.txt`` (reference training.py:398-440)."""
    args = get_args()
    if not getattr(args, "log_progress", False) or args.save is None:
        return
    if barrier and dist.is_initialized():
        dist.barrier()
    if not dist.is_initialized() or dist.get_rank() == 0:
        os.makedirs(args.save, exist_ok=True)
        job = os.getenv("SLURM_JOB_ID", "") or os.getenv("TORCHELASTIC_RUN_ID", "") or str(os.getpid())
        with open(os.path.join(args.save, "progress.txt"), "a") as f:
            f.write(f"{time.strftime('%Y-%m-%d %H:%M:%S')}\tJob ID: {job}\t# GPUs: {args.world_size}\t{line}\n")


def compute_throughputs_and_append_to_progress_log(iteration, num_floating_point_operations_so_far):
    """Job throughput (since this launch) in TFLOP/s/GPU from the analytic FLOP model, logged at every checkpoint."""
    args = get_args()
    if not getattr(args, "log_progress", False) or args.save is None:
        return
    done = num_floating_point_operations_so_far - getattr(args, "_flops_at_job_start", 0)
    tput = done / (max(time.time() - _TRAIN_START_TIME, 1e-9) * 1e12 * args.world_size)
    tokens = args.consumed_train_samples * args.seq_length
    append_to_progress_log(f"{'Saving async checkpoint' if args.async_save else 'Saved checkpoint'}\tIteration: {iteration}\t"
                           f"Job throughput: {tput:.1f} TFLOP/s/GPU\t"
                           f"Floating-point operations: {num_floating_point_operations_so_far:.2e}\t"
                           f"Tokens (in billions): {tokens / 1e9:.2f}")


def save_checkpoint_and_time(iteration, model, optimizer, opt_param_scheduler, num_floating_point_operations_so_far):
    timers = get_timers()
    timers("save-checkpoint", log_level=0).start(barrier=True)
    ft_integration.on_checkpointing_start()
    checkpointing.save_checkpoint(iteration, model, optimizer, opt_param_scheduler, num_floating_point_operations_so_far)
    ft_integration.on_checkpointing_end(is_async_finalization=False)
    timers("save-checkpoint").stop(barrier=True)
    timers.log(["save-checkpoint"])
    compute_throughputs_and_append_to_progress_log(iteration, num_floating_point_operations_so_far)


def train(forward_step_func, model, optimizer, opt_param_scheduler, train_data_iterator, valid_data_iterator,
          process_non_loss_data_func, config):
    args = get_args()
    timers = get_timers()
    tracer = get_tracer()
    for m in model:
        m.train()
    total_loss_dict = {}
    iteration = args.iteration
    nflops = args.num_floating_point_operations_so_far
    config.grad_scale_func = optimizer.scale_loss
    config.timers = timers
    config.finalize_model_grads_func = finalize_model_grads
    timers("interval-time", log_level=0).start(barrier=True)
    report_memory_flag = True
    ws_server = None
    if args.training_ws_port is not None:
        from ..scope.training_server import maybe_start_training_server

        ws_server = maybe_start_training_server(args, model)
    slow = None
    if args.slow_rank_inject:
        r, cyc = args.slow_rank_inject.split(":")
        if int(r) == args.rank:
            slow = int(cyc)
            from .. import ops

            ops.set_slow_rank_spin(slow)      # spins in front of every layer, forward and backward
    if args.async_save and getattr(args, "use_persistent_ckpt_worker", False) and args.ckpt_format == "torch_dist":
        from .async_utils import init_persistent_async_worker

        init_persistent_async_worker()        # spawn the writer process now, not inside the first save
    exit_flag = False
    start_time = time.time()
    stimer = get_straggler_detector()
    if args.log_straggler:
        stimer.configure(args.world_size, args.rank, mmcnt=args.straggler_minmax_count,
                         enabled=not args.disable_straggler_on_startup, port=args.straggler_ctrlr_port)
    if args.manual_gc:
        gc.disable()
        gc.collect()
    prof = _Profiling(args)
    while iteration < args.train_iters:
        prof.step_begin(iteration)
        if ws_server is not None:
            ws_server.wait_for_step_request(model)
        tracer.iteration_begin(iteration)
        update_num_microbatches(args.consumed_train_samples, consistency_check=True)
        if iteration in args.iterations_to_skip:
            for _ in range(get_num_microbatches()):
                next(train_data_iterator) if train_data_iterator is not None else None
            iteration += 1
            args.consumed_train_samples += get_current_global_batch_size()
            tracer.iteration_end()
            continue
        args.curr_iteration = iteration
        _maybe_simulate_fault(args, iteration)
        if getattr(args, '_fault_deadline', None) and time.time() > args._fault_deadline:
            while True:
                time.sleep(3600)
        ft_integration.on_training_step_start()
        with stimer:
            loss_dict, skipped_iter, should_ckpt, should_exit, exit_code, grad_norm, num_zeros = train_step(
                forward_step_func, train_data_iterator, model, optimizer, opt_param_scheduler, config)
        ft_integration.on_training_step_end()
        if should_exit:
            if should_ckpt and args.save:
                save_checkpoint_and_time(iteration, model, optimizer, opt_param_scheduler, nflops)
            print_rank_0(f"rerun state machine requested exit (code {exit_code})")
            args.rerun_exit_code = exit_code
            exit_flag = True
            break
        iteration += 1
        batch_size = ps.get_data_parallel_world_size() * args.micro_batch_size * get_num_microbatches()
        args.consumed_train_samples += batch_size
        nflops += num_floating_point_operations(args, batch_size)
        tracer.iteration_end()
        loss_scale = float(optimizer.get_loss_scale().item()) if (args.fp16 or args.loss_scale) else 1.0
        params_norm = calc_params_l2_norm(model) if args.log_params_norm else None
        lr = None
        for g in optimizer.param_groups:
            if g["lr"] is not None:
                lr = g["lr"]
                break
        report_memory_flag = training_log(loss_dict, total_loss_dict, lr or 0.0, iteration, loss_scale,
                                          report_memory_flag, skipped_iter, grad_norm, params_norm, num_zeros)
        prof.step_end(iteration)
        if args.check_weight_hash_across_dp_replicas_interval and \
                iteration % args.check_weight_hash_across_dp_replicas_interval == 0:
            assert check_param_hashes_across_dp_replicas(model), "parameter hashes differ across data-parallel replicas"
            print_rank_0(f">>> weight hashes match across DP replicas at iteration {iteration}")
        if args.log_straggler and iteration % args.log_interval == 0:
            stimer.report(num_floating_point_operations(args, batch_size) * args.log_interval, args.log_interval)
        if args.eval_interval and iteration % args.eval_interval == 0 and getattr(args, "do_valid", valid_data_iterator is not None):
            timers("interval-time").stop()
            evaluate_and_print_results(f"iteration {iteration}", forward_step_func, valid_data_iterator, model,
                                       iteration, config)
            timers("interval-time", log_level=0).start(barrier=True)
        if args.manual_gc and args.manual_gc_interval and iteration % args.manual_gc_interval == 0:
            gc.collect()
        if getattr(args, "train_sync_interval", None) and iteration % args.train_sync_interval == 0 \
                and torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()        # keep the host from running arbitrarily far ahead of the GPU
        checkpointing.maybe_finalize_async_save(blocking=False)
        saved = False
        sh = get_signal_handler()
        if sh is not None and any(sh.signals_received()):
            if args.save:
                save_checkpoint_and_time(iteration, model, optimizer, opt_param_scheduler, nflops)
            print_rank_0("exiting program after receiving SIGTERM.")
            exit_flag = True
            break
        if getattr(args, "adlr_autoresume", False) and iteration % args.adlr_autoresume_interval == 0:
            from .autoresume import check_adlr_autoresume_termination

            def _save(it, m, o, s_):
                if args.save:
                    save_checkpoint_and_time(it, m, o, s_, nflops)

            if check_adlr_autoresume_termination(iteration, model, optimizer, opt_param_scheduler, _save):
                exit_flag = True
                break
        if (args.save and getattr(args, "non_persistent_save_interval", None) and not saved
                and iteration % args.non_persistent_save_interval == 0
                and not (args.save_interval and iteration % args.save_interval == 0)):
            if args.non_persistent_ckpt_type not in ("global", "local"):
                raise NotImplementedError("--non-persistent-ckpt-type must be global or local")
            timers("interval-time").stop()
            checkpointing.save_checkpoint(iteration, model, optimizer, opt_param_scheduler, nflops, non_persistent_ckpt=True)
            timers("interval-time", log_level=0).start(barrier=True)
        if args.save and args.save_interval and iteration % args.save_interval == 0:
            timers("interval-time").stop()
            save_checkpoint_and_time(iteration, model, optimizer, opt_param_scheduler, nflops)
            saved = True
            timers("interval-time", log_level=0).start(barrier=True)
        if args.exit_duration_in_mins is not None:
            mins = (time.time() - _TRAIN_START_TIME) / 60.0
            flag = torch.tensor([int(mins > args.exit_duration_in_mins)], dtype=torch.int,
                                device="cuda" if (torch.cuda.is_available() and args.distributed_backend == "nccl") else "cpu")
            if dist.is_initialized():
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if flag.item():
                if args.save and not saved:
                    save_checkpoint_and_time(iteration, model, optimizer, opt_param_scheduler, nflops)
                print_rank_0(f"exiting program after {mins} minutes")
                exit_flag = True
                break
        if args.exit_interval and iteration % args.exit_interval == 0:
            if args.save and not saved:
                save_checkpoint_and_time(iteration, model, optimizer, opt_param_scheduler, nflops)
            print_rank_0(f"exiting program at iteration {iteration}")
            exit_flag = True
            break
    tracer.iteration_end()
    checkpointing.maybe_finalize_async_save(blocking=True)
    if ws_server is not None:
        ws_server.shutdown()
    args.num_floating_point_operations_so_far = nflops
    return iteration, nflops, exit_flag


class _Profiling:
    """``--profile`` (CUDA profiler range for nsys/ncu capture windows), ``--use-pytorch-profiler`` (Kineto
    trace to ``--tensorboard-dir``), ``--record-memory-history`` (allocator snapshot dumped at the end of the
    window or on OOM) -- reference training.py:2134-2177."""

    def __init__(self, args):
        self.args = args
        self.on = bool(args.profile) and args.rank in (args.profile_ranks or [0])
        self.torch_prof = None
        self.active = False
        if getattr(args, "record_memory_history", False) and torch.cuda.is_available():
            torch.cuda.memory._record_memory_history(max_entries=100000)
            self.mem = True
        else:
            self.mem = False

    def step_begin(self, iteration):
        a = self.args
        if not self.on or iteration != a.profile_step_start or self.active:
            return
        self.active = True
        if a.use_pytorch_profiler:
            import torch.profiler as tp

            acts = [tp.ProfilerActivity.CPU] + ([tp.ProfilerActivity.CUDA] if torch.cuda.is_available() else [])
            out = a.tensorboard_dir or "."
            self.torch_prof = tp.profile(activities=acts, record_shapes=True, with_stack=False,
                                         on_trace_ready=tp.tensorboard_trace_handler(out))
            self.torch_prof.start()
        elif torch.cuda.is_available():
            torch.cuda.cudart().cudaProfilerStart()

    def step_end(self, iteration):
        a = self.args
        if self.active and self.torch_prof is not None:
            self.torch_prof.step()
        if not self.active or iteration < a.profile_step_end:
            return
        self.active = False
        if self.torch_prof is not None:
            self.torch_prof.stop()
            self.torch_prof = None
        elif torch.cuda.is_available():
            torch.cuda.cudart().cudaProfilerStop()
        if self.mem:
            torch.cuda.memory._dump_snapshot(a.memory_snapshot_path)
            print_rank_0(f"memory snapshot written to {a.memory_snapshot_path}")


def check_param_hashes_across_dp_replicas(model) -> bool:
    """Every data-parallel replica must hold bit-identical parameters (reference
    ``core/utils.py check_param_hashes_across_dp_replicas``): compare a cheap order-sensitive fp64 checksum."""
    g = ps.get_data_parallel_group()
    if g is None or isinstance(g, ps._SingleGroup) or dist.get_world_size(g) == 1:
        return True
    sums = []
    for m in model:
        for p in m.parameters():
            d = p.detach().double().flatten()
            idx = torch.arange(1, d.numel() + 1, device=d.device, dtype=torch.float64)
            sums.append((d * (idx % 1021 + 1)).sum())
    local = torch.stack(sums) if sums else torch.zeros(1, dtype=torch.float64)
    dev = "cuda" if (torch.cuda.is_available() and dist.get_backend(g) == "nccl") else "cpu"
    local = local.to(dev)
    gathered = [torch.zeros_like(local) for _ in range(dist.get_world_size(g))]
    dist.all_gather(gathered, local, group=g)
    return all(torch.equal(gathered[0], x) for x in gathered[1:])


def _maybe_simulate_fault(args, iteration):
    """``--simulated-fault kind:rank:delay_s`` (reference training.py ``--simulated-fault``): hang or
    kill one rank after a delay so failure detection / restart logic can be exercised."""
    spec = getattr(args, "simulated_fault", None)
    if not spec or getattr(args, "_fault_armed", False):
        return
    kind, rank, delay = (spec.replace(";", ":").split(":") + ["0", "0"])[:3]
    if int(rank) != args.rank:
        return
    args._fault_armed = True
    import threading

    def fire():
        time.sleep(float(delay))
        if kind == "rank_killed":
            os.kill(os.getpid(), 9)
        else:   # rank_hung
            while True:
                time.sleep(3600)

    if kind == "rank_killed":
        threading.Thread(target=fire, daemon=True).start()
    else:
        # a hung rank blocks its main thread at the next iteration boundary after the delay
        args._fault_deadline = time.time() + float(delay)


def cyclic_iter(it):
    while True:
        for x in it:
            yield x


def build_train_valid_test_data_iterators(provider):
    args = get_args()
    from ..data.samplers import build_pretraining_data_loader

    train_samples = args.train_iters * args.global_batch_size if args.train_iters else args.train_samples
    eval_iters = (args.train_iters // args.eval_interval + 1) * args.eval_iters if args.eval_interval else args.eval_iters
    nums = [train_samples, eval_iters * args.global_batch_size, args.eval_iters * args.global_batch_size]
    is_dist = getattr(provider, "is_distributed", False)
    need = is_dist or ps.get_tensor_model_parallel_rank() == 0
    iters = [None, None, None]
    if need:
        train_ds, valid_ds, test_ds = provider(nums)
        dp_rank = ps.get_data_parallel_rank()
        dp_size = ps.get_data_parallel_world_size()
        for i, (ds, consumed) in enumerate(((train_ds, args.consumed_train_samples),
                                            (valid_ds, args.consumed_valid_samples), (test_ds, 0))):
            dl = build_pretraining_data_loader(ds, consumed, args.micro_batch_size, dp_rank, dp_size,
                                               args.dataloader_type, args.num_workers, args.seed)
            if dl is not None:
                iters[i] = iter(dl) if args.dataloader_type == "single" else iter(cyclic_iter(dl))
    # every rank must agree on whether a split exists (only TP rank 0 holds iterators): reference
    # training.py build_train_valid_test_data_loaders broadcasts do_train / do_valid / do_test the same way
    flags = torch.tensor([int(it is not None) for it in iters], dtype=torch.long,
                         device="cuda" if (torch.cuda.is_available() and args.distributed_backend == "nccl") else "cpu")
    if dist.is_initialized():
        dist.all_reduce(flags, op=dist.ReduceOp.MAX)
    args.do_train = bool(flags[0]) or bool(getattr(args, "do_train", False))
    args.do_valid = bool(flags[1]) or bool(getattr(args, "do_valid", False))
    args.do_test = bool(flags[2]) or bool(getattr(args, "do_test", False))
    return tuple(iters)


def pretrain(train_valid_test_dataset_provider, model_provider, model_type, forward_step_func,
             process_non_loss_data_func=None, extra_args_provider=None, args_defaults=None,
             get_embedding_ranks=None, get_position_embedding_ranks=None, non_loss_data_func=None, args_list=None):
    initialize_megatron(extra_args_provider=extra_args_provider,
                        args_defaults={**(args_defaults or {}), "model_type_hint": str(model_type)},
                        args_list=args_list)
    args = get_args()
    timers = get_timers()
    print_rank_0(f"time to initialize megatron (seconds): {time.time() - _TRAIN_START_TIME:.3f}")
    if args.enable_ft_package:
        ft_integration.setup(args)
        ft_integration.maybe_setup_simulated_fault()
    if args.use_dpp:
        from ..dpp.scheduler import install_dpp_schedule

        install_dpp_schedule(args)
    if args.forward_backward_disaggregating:
        from ..fbd.schedule import install_fbd_schedule

        install_fbd_schedule(args)
    timers("model-and-optimizer-setup", log_level=0).start(barrier=True)
    append_to_progress_log("Starting job")
    model, optimizer, scheduler = setup_model_and_optimizer(model_provider, model_type)
    if args.forward_backward_disaggregating:
        # forward ranks start from their backward twins' weights (mandatory with --ignore-forward-tensor-parallel, where
        # the un-sharded forward rank cannot reproduce the TP-sharded initialisation from the seed)
        from ..fbd.schedule import sync_forward_weights

        sync_forward_weights(model)
    timers("model-and-optimizer-setup").stop()
    args._flops_at_job_start = args.num_floating_point_operations_so_far
    if getattr(args, "ckpt_convert_format", None):
        # --ckpt-convert-format F --ckpt-convert-save DIR: re-save the just-loaded checkpoint in another format and stop
        # (reference training.py: "Convert checkpoint format" block)
        if not args.ckpt_convert_save:
            raise ValueError("--ckpt-convert-format needs --ckpt-convert-save")
        args.ckpt_format, args.save = args.ckpt_convert_format, args.ckpt_convert_save
        checkpointing.save_checkpoint(args.iteration, model, optimizer, scheduler, args.num_floating_point_operations_so_far)
        print_rank_0(f"> converted checkpoint: {args.load} -> {args.save} ({args.ckpt_format})")
        return model, optimizer
    config = unwrap_model(model[0]).config
    timers("train/valid/test-data-iterators-setup", log_level=0).start(barrier=True)
    if args.virtual_pipeline_model_parallel_size is not None:
        tr, va, te = [], [], []
        for i in range(len(model)):
            ps.set_virtual_pipeline_model_parallel_rank(i)
            a, b, c = build_train_valid_test_data_iterators(train_valid_test_dataset_provider)
            tr.append(a); va.append(b); te.append(c)
        ps.set_virtual_pipeline_model_parallel_rank(0)
        train_it, valid_it, test_it = tr, va, te
    else:
        train_it, valid_it, test_it = build_train_valid_test_data_iterators(train_valid_test_dataset_provider)
    timers("train/valid/test-data-iterators-setup").stop()
    print_rank_0("done with setup ...")
    timers.log(["model-and-optimizer-setup", "train/valid/test-data-iterators-setup"], barrier=True)
    iteration, nflops = args.iteration, args.num_floating_point_operations_so_far
    initialize_rerun_state_machine(
        mode=args.rerun_mode,
        error_injector=RerunErrorInjector(args.error_injection_rate, args.error_injection_type, seed=args.seed))
    if args.rerun_mode != "disabled":
        wrap = lambda it: RerunDataIterator(it) if it is not None else None
        train_it = [wrap(i) for i in train_it] if isinstance(train_it, list) else wrap(train_it)
    if not args.skip_train and args.dataloader_type == "cyclic" and getattr(args, "retro_project_dir", None):
        # reference training.py:806-809: cyclic Retro training runs for --retro-cyclic-train-iters
        if args.retro_cyclic_train_iters is None:
            raise ValueError("--dataloader-type cyclic with a retro project needs --retro-cyclic-train-iters")
        args.train_iters = args.retro_cyclic_train_iters
        print_rank_0(f"retro cyclic train iters : {args.train_iters}")
    if not args.skip_train and args.train_iters and args.train_iters > 0:
        print_rank_0("training ...")
        iteration, nflops, exited = train(forward_step_func, model, optimizer, scheduler, train_it, valid_it,
                                          process_non_loss_data_func, config)
        print_rank_0(f"after training is done")
        if args.save and iteration != 0 and (args.save_interval is None or iteration % args.save_interval != 0) and not exited:
            checkpointing.save_checkpoint(iteration, model, optimizer, scheduler, nflops)
    tracer = get_tracer()
    if tracer.enabled:
        tracer.shutdown()
    # final evaluation on the validation and the test split -- also the whole job of an evaluation-only run (--skip-train)
    if args.eval_iters and getattr(args, "do_valid", valid_it is not None) and (args.eval_interval or args.skip_train):
        evaluate_and_print_results(f"the end of training for val data", forward_step_func, valid_it, model,
                                   iteration, config)
    if args.eval_iters and getattr(args, "do_test", False) and (args.eval_interval or args.skip_train):
        evaluate_and_print_results(f"the end of training for test data", forward_step_func, test_it, model,
                                   iteration, config)
    if args.use_dpp:
        from ..dpp.scheduler import shutdown_dpp

        shutdown_dpp()
    p2p.flush()
    ft_integration.shutdown()
    return model, optimizer


# -- further names of the reference's ``training.py`` ------------------------------------------------------------------------
def print_datetime(string: str):
    if dist.is_initialized():
        dist.barrier()
    print_rank_0(f"[{string}] datetime: {time.strftime('%Y-%m-%d %H:%M:%S')} ")


def destroy_global_state():
    """Tear everything down so that ``pretrain`` can be called again in the same process (tests, notebooks)."""
    from .global_vars import unset_global_variables

    unset_global_variables()
    p2p.destroy_channels()
    ps.destroy_model_parallel()


def update_train_iters(args):
    """``--train-samples`` -> ``args.train_iters`` (constant batch size, or with the batch-size ramp-up accounted for)."""
    if args.train_iters:
        return
    if args.rampup_batch_size is None:
        args.train_iters = args.train_samples // args.global_batch_size
    else:
        start, inc, ramp = (int(x) for x in args.rampup_batch_size)
        iters, consumed, bs = 0, 0, start
        while consumed <= ramp and consumed < args.train_samples:
            bs = min(args.global_batch_size, start + (consumed * (args.global_batch_size - start) // max(ramp, 1)) // inc * inc)
            consumed += bs
            iters += 1
        iters += max(0, args.train_samples - consumed) // args.global_batch_size
        args.train_iters = iters
    print_rank_0(f"setting training iterations to {args.train_iters}")


def get_start_time_from_progress_log():
    """(start time, floating-point operations so far) of the last job in ``<save>/progress.txt`` that shares this job's world
    size -- what the throughput summary after a restart is computed against."""
    from datetime import datetime

    args = get_args()
    path = os.path.join(args.save, "progress.txt")
    start_time = start_flops = None
    latest_flops = 0.0
    with open(path) as f:
        for line in f:
            parts = line.strip().split("\t")
            ws = [p for p in parts if p.startswith("# GPUs: ")]
            if ws and int(ws[0].split(": ")[1]) != args.world_size:
                start_time = start_flops = None                       # another world size: the series restarts
                continue
            if "Saved checkpoint" in line:
                fl = [p for p in parts if p.startswith("Floating-point operations: ")]
                if fl:
                    latest_flops = float(fl[0].split(": ")[1])
            if "Starting job" in line and start_time is None:
                start_time = datetime.strptime(parts[0], "%Y-%m-%d %H:%M:%S")
                start_flops = latest_flops
    if start_time is None:
        raise RuntimeError(f"no job start found in {path}")
    return start_time, start_flops


def get_train_valid_test_num_samples():
    """Samples the train / validation / test datasets must provide for this run (reference ``training.py``)."""
    args = get_args()
    train_samples = args.train_samples if args.train_samples else args.train_iters * args.global_batch_size
    eval_iters = (args.train_iters // args.eval_interval + 1) * args.eval_iters if args.eval_interval else args.eval_iters
    return train_samples, eval_iters * args.global_batch_size, args.eval_iters * args.global_batch_size


def enable_forward_pre_hook(model_chunks):
    """Turn the overlapped parameter all-gather back on (after evaluation / checkpointing ran with it off)."""
    for m in model_chunks:
        if hasattr(m, "enable_forward_pre_hook"):
            m.enable_forward_pre_hook()


def disable_forward_pre_hook(model_chunks, param_sync: bool = True):
    """Turn the overlapped parameter all-gather off; with ``param_sync`` the parameters are synchronised once first."""
    for m in model_chunks:
        if hasattr(m, "disable_forward_pre_hook"):
            m.disable_forward_pre_hook(param_sync=param_sync)
        elif param_sync and hasattr(m, "start_param_sync"):
            m.start_param_sync(force_sync=True)


def dummy_train_step(data_iterator):
    """Consume one global batch without computing anything (``--iterations-to-skip``)."""
    from ..utils.microbatches import get_num_microbatches as _n

    for _ in range(_n()):
        next(data_iterator) if not isinstance(data_iterator, list) else [next(it) for it in data_iterator]


def preprocess_common_state_dict(common_state_dict):
    """What ranks must agree on in the common part of a distributed checkpoint: drop the per-rank entries of ``args``."""
    import copy

    sd = copy.deepcopy(common_state_dict)
    a = sd.get("args")
    if a is not None:
        d = vars(a) if not isinstance(a, dict) else a
        for k in ("rank", "local_rank"):
            d.pop(k, None)
        sd["args"] = d
    return sd
