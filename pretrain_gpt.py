"""Pretrain GPT (entry-point name and provider protocol kept from the reference
``pretrain_gpt.py``: ``model_provider(pre_process, post_process)``, ``forward_step(data_iterator,
model) -> (output_tensor, partial(loss_func, loss_mask))``, ``loss_func -> (loss, num_tokens,
{'lm loss': (sum, count)})``, ``train_valid_test_datasets_provider(train_val_test_num_samples)``)."""
import os

if not os.path.exists("/dev/nvidia0"):   # CPU/gloo mode: avoid OpenMP oversubscription
    os.environ.setdefault("OMP_NUM_THREADS", "2")

import math
from functools import partial

import torch

from megatronapp_b200.data.mock import MockGPTDataset
from megatronapp_b200.models.gpt import GPTModel, get_gpt_decoder_block_spec, get_gpt_layer_local_spec
from megatronapp_b200.parallel import state as mpu
from megatronapp_b200.training.arguments import core_transformer_config_from_args
from megatronapp_b200.training.global_vars import get_args, get_timers
from megatronapp_b200.training.rerun_state_machine import get_rerun_state_machine
from megatronapp_b200.training.training import pretrain
from megatronapp_b200.training.utils import get_batch_on_this_tp_rank, print_rank_0


SPIKY_LOSS_FACTOR = 10          # a loss 10x the largest seen so far is re-examined (reference pretrain_gpt.py)


def model_provider(pre_process=True, post_process=True):
    args = get_args()
    student = _build_model(args, pre_process, post_process)
    if getattr(args, "export_kd_teacher_load", None) and not args.export_kd_finalize:
        _attach_distillation_teacher(args, student)
    return student


def _attach_distillation_teacher(args, student):
    """--export-kd-teacher-load: knowledge distillation from a frozen teacher checkpoint (reference
    ``megatron/post_training/model_provider.py:222-247``).  The teacher shares the student's tensor-parallel layout and its
    tokenizer; its architecture comes from ``model_config.yaml`` in the checkpoint folder or from the checkpoint's own args."""
    from megatronapp_b200.post_training import attach_teacher, load_distillation_config, teacher_args_from_checkpoint
    from megatronapp_b200.training.checkpointing import load_checkpoint

    if args.pipeline_model_parallel_size > 1:
        raise ValueError("distillation does not support pipeline parallelism")
    if args.use_legacy_models:
        raise ValueError("distillation needs the core GPT model (not --use-legacy-models)")
    targs = teacher_args_from_checkpoint(args, args.export_kd_teacher_load)
    targs.export_kd_teacher_load, targs.export_kd_finalize = None, False
    print_rank_0(f"building the distillation teacher ({targs.num_layers} layers, hidden size {targs.hidden_size}) ...")
    teacher = _build_model(targs, True, True)
    finetune, no_load_rng = args.finetune, args.no_load_rng
    args.finetune = args.no_load_rng = True          # weights only: no iteration / rng / consumed-samples side effects
    try:
        load_checkpoint([teacher], None, None, load_arg="export_kd_teacher_load", strict=True)
    finally:
        args.finetune, args.no_load_rng = finetune, no_load_rng
    if args.bf16 or args.fp16:
        teacher = teacher.to(torch.bfloat16 if args.bf16 else torch.float16)
    if torch.cuda.is_available():
        teacher = teacher.cuda(torch.cuda.current_device())
    cfg = load_distillation_config(args.export_kd_cfg, student.config, teacher.config, args.sequence_parallel)
    attach_teacher(student, teacher, cfg, args.sequence_parallel)


def _build_model(args, pre_process=True, post_process=True):
    print_rank_0("building GPT model ...")
    config = core_transformer_config_from_args(args)
    if args.use_legacy_models:
        from megatronapp_b200.legacy.model import GPTModel as LegacyGPTModel

        return LegacyGPTModel(config, num_tokentypes=0, parallel_output=True, pre_process=pre_process,
                              post_process=post_process)
    if args.spec:
        # --spec module.path [function]: user-supplied layer / block spec (reference pretrain_gpt.py:96-97)
        from megatronapp_b200.transformer.spec_utils import import_module

        obj = import_module(tuple(args.spec) if len(args.spec) > 1 else tuple(args.spec[0].rsplit(".", 1)))
        spec = obj
        if callable(obj) and not hasattr(obj, "module"):            # a factory rather than a ModuleSpec instance
            try:
                spec = obj()
            except TypeError:
                spec = obj(config)
        return GPTModel(config=config, transformer_layer_spec=spec, vocab_size=args.padded_vocab_size,
                        max_sequence_length=args.max_position_embeddings, pre_process=pre_process, post_process=post_process,
                        fp16_lm_cross_entropy=args.fp16_lm_cross_entropy, parallel_output=True,
                        share_embeddings_and_output_weights=not args.untie_embeddings_and_output_weights,
                        position_embedding_type=args.position_embedding_type, rotary_percent=args.rotary_percent,
                        rotary_base=args.rotary_base)
    het = getattr(args, "heterogeneous_layers_config_path", None) or getattr(args, "heterogeneous_layers_config_encoded_json", None)
    if het:
        import json

        from megatronapp_b200.models.gpt.heterogeneous.heterogeneous_layer_specs import (
            get_gpt_heterogeneous_layer_spec, parse_heterogeneous_config)

        cfg = het if args.heterogeneous_layers_config_path else json.loads(het)
        spec = get_gpt_heterogeneous_layer_spec(config, parse_heterogeneous_config(cfg, config.num_attention_heads,
                                                                                   config.hidden_size))
    elif args.num_experts:
        spec = get_gpt_decoder_block_spec(config)
    else:
        spec = get_gpt_layer_local_spec(args.num_experts, args.moe_grouped_gemm, args.qk_layernorm,
                                        args.multi_latent_attention)
    return GPTModel(config=config, transformer_layer_spec=spec, vocab_size=args.padded_vocab_size,
                    max_sequence_length=args.max_position_embeddings, pre_process=pre_process,
                    post_process=post_process, fp16_lm_cross_entropy=args.fp16_lm_cross_entropy,
                    parallel_output=True,
                    share_embeddings_and_output_weights=not args.untie_embeddings_and_output_weights,
                    position_embedding_type=args.position_embedding_type, rotary_percent=args.rotary_percent,
                    rotary_base=args.rotary_base, rope_scaling=args.use_rope_scaling,
                    rope_scaling_factor=args.rope_scaling_factor,
                    seq_len_interpolation_factor=args.rotary_seq_len_interpolation_factor)


def get_batch(data_iterator):
    args = get_args()
    if (not mpu.is_pipeline_first_stage()) and (not mpu.is_pipeline_last_stage()):
        return None, None, None, None, None
    b = get_batch_on_this_tp_rank(data_iterator, args)
    if args.context_parallel_size > 1:
        from megatronapp_b200.parallel.context_parallel import get_batch_on_this_cp_rank

        b = get_batch_on_this_cp_rank(b)
    return b["tokens"], b["labels"], b["loss_mask"], b["attention_mask"], b["position_ids"]


def loss_func(loss_mask, output_tensor):
    args = get_args()
    losses = output_tensor.float()
    loss_mask = loss_mask.view(-1).float()
    total_tokens = loss_mask.sum()
    loss = torch.cat([torch.sum(losses.view(-1) * loss_mask).view(1), total_tokens.view(1)])
    if args.context_parallel_size > 1:
        torch.distributed.all_reduce(loss, group=mpu.get_context_parallel_group())
    rsm = get_rerun_state_machine()
    if rsm.mode != "disabled":
        # NaN / Inf checks with re-execution (reference pretrain_gpt.py loss_func).  With reruns disabled the reference
        # still synchronises on every microbatch loss here; this framework leaves that to the per-iteration NaN
        # accounting of the training loop.
        rsm.validate_result(loss[0] / loss[1].clamp(min=1), lambda v: not math.isfinite(v), "found NaN/Inf in local loss")
    if args.check_for_spiky_loss:
        rsm.validate_result(loss[0] / loss[1].clamp(min=1),
                            partial(rsm.is_unexpectedly_large, threshold=SPIKY_LOSS_FACTOR, context="loss"), "Spiky loss",
                            tolerance=0.0, fatal=False)
    reporting = loss.clone().detach()
    g = mpu.get_data_parallel_group()
    if g is not None and not isinstance(g, mpu._SingleGroup) and torch.distributed.get_world_size(g) > 1:
        torch.distributed.all_reduce(reporting, group=g)
    local_num_tokens = loss[1].clone().detach().to(torch.int)
    return loss[0] * args.context_parallel_size, local_num_tokens, {"lm loss": (reporting[0], reporting[1])}


def kd_loss_func(loss_mask, kd, output_tensor):
    """Loss with the distillation terms (reference ``post_training/loss_func.py``): the LM loss is still reported, the
    optimised quantity is the balancer's mix of LM, logit-KL and intermediate-layer losses."""
    lm_sum, num_tokens, report = loss_func(loss_mask, output_tensor)
    mask = loss_mask.float()
    count = mask.sum().clamp(min=1)
    total, kd_logits = kd.compute_kd_loss(student_loss=lm_sum / count, loss_reduction_fn=lambda x: (x * mask).sum() / count)
    rep = torch.stack([kd_logits * count, count]).detach()
    g = mpu.get_data_parallel_group()
    if g is not None and not isinstance(g, mpu._SingleGroup) and torch.distributed.get_world_size(g) > 1:
        torch.distributed.all_reduce(rep, group=g)
    report["kd loss"] = (rep[0], rep[1])
    return total * count, num_tokens, report


def forward_step(data_iterator, model):
    timers = get_timers()
    timers("batch-generator", log_level=2).start()
    tokens, labels, loss_mask, attention_mask, position_ids = get_batch(data_iterator)
    timers("batch-generator").stop()
    output_tensor = model(tokens, position_ids, attention_mask, labels=labels)
    core = model
    while hasattr(core, "module"):
        core = core.module
    kd = core.__dict__.get("_kd")
    if kd is not None and core.training:
        kd.teacher_forward(tokens, position_ids, attention_mask)
        return output_tensor, partial(kd_loss_func, loss_mask, kd)
    return output_tensor, partial(loss_func, loss_mask)


def train_valid_test_datasets_provider(train_val_test_num_samples):
    args = get_args()
    print_rank_0("> building train, validation, and test datasets for GPT ...")
    real_data = (args.data_path or args.data_args_path or args.per_split_data_args_path or args.train_data_path
                 or args.valid_data_path or args.test_data_path)
    if args.mock_data or not real_data:
        vs = args.padded_vocab_size
        mk = lambda n, seed: MockGPTDataset(max(int(n), 1), args.seq_length, vs, seed,  # noqa: E731
                                            create_attention_mask=args.create_attention_mask_in_dataloader)
        return mk(train_val_test_num_samples[0], args.seed), mk(train_val_test_num_samples[1], args.seed + 1), \
            mk(train_val_test_num_samples[2], args.seed + 2)
    from megatronapp_b200.data.gpt_dataset import build_gpt_datasets

    return build_gpt_datasets(args, train_val_test_num_samples)


train_valid_test_datasets_provider.is_distributed = False

if __name__ == "__main__":
    pretrain(train_valid_test_datasets_provider, model_provider, "encoder_or_decoder", forward_step,
             args_defaults={"tokenizer_type": "NullTokenizer"})
