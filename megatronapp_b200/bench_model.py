"""Public "few lines to a training step" API used by ``bench.py``, ``smoke()`` and the examples.

``build_gpt_for_bench`` sets up (single- or multi-rank) parallel state, a GPT model wrapped in
DDP + the fused-Adam optimizer; ``train_step`` runs forward/backward over a list of microbatches
through the regular schedule and one optimizer step, returning the (host) loss.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from functools import partial
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from .config.transformer_config import TransformerConfig
from .models.gpt import GPTModel, get_gpt_layer_local_spec
from .optim import OptimizerConfig, get_megatron_optimizer
from .parallel import p2p, schedules
from .parallel import state as ps
from .parallel.ddp import DistributedDataParallel, DistributedDataParallelConfig, finalize_model_grads
from .parallel.random import model_parallel_cuda_manual_seed
from .transformer.module import Float16Module


@dataclass
class BenchState:
    model: List
    optimizer: object
    config: TransformerConfig
    seq: int
    micro_batch: int
    vocab: int
    pre_process: bool
    post_process: bool
    extras: Dict = field(default_factory=dict)


def build_gpt_for_bench(num_layers, hidden, heads, seq, vocab, micro_batch, ffn=None, tp=1, pp=1,
                        sequence_parallel=False, device="cuda", dtype=torch.bfloat16, swiglu=False,
                        rope=False, rmsnorm=False, kv_groups=None, bias=True, lr=1e-4, fused_tp_comm=True,
                        use_distributed_optimizer=False, dropout=0.0, vpp=None, overlap_grad_reduce=False,
                        seed=1234, num_experts=None, ep=1, moe_topk=2, fused_grad_rs=False, fp8=None, fp8_recipe="delayed",
                        moe_grouped=False, moe_dispatcher=None, moe_ffn=None, fbd=False, etp=None,
                        moe_aux_loss_coeff=0.0, fbd_collapse=False, recompute_layers=0, bf16_adam_moments=False) -> BenchState:
    if torch.cuda.is_available() and str(device).startswith("cuda"):
        torch.cuda.set_device(torch.device(device) if ":" in str(device) else torch.cuda.current_device())
    if not ps.model_parallel_is_initialized():
        kw_ps = dict(virtual_pipeline_model_parallel_size=vpp, expert_model_parallel_size=ep)
        if etp is not None:
            kw_ps["expert_tensor_parallel_size"] = etp
        if fbd:
            kw_ps["forward_backward_disaggregating"] = True
            kw_ps["ignore_forward_tensor_parallel"] = bool(fbd_collapse)
        ps.initialize_model_parallel(tp, pp, **kw_ps)
        if dist.is_initialized():
            p2p.build_channels()
    model_parallel_cuda_manual_seed(seed)
    fbd_bwd = None
    if ps.fbd_forward_tp_collapsed() and ps.is_forward_stage():
        import types

        fbd_bwd = types.SimpleNamespace(sequence_parallel=bool(sequence_parallel and tp > 1), tensor_model_parallel_size=tp)
        tp, sequence_parallel = 1, False          # the un-sharded forward rank runs true TP=1 math
    kw = dict(fbd_backward_config=fbd_bwd, num_layers=num_layers, hidden_size=hidden, num_attention_heads=heads, ffn_hidden_size=ffn,
              num_query_groups=kv_groups, tensor_model_parallel_size=tp, pipeline_model_parallel_size=pp,
              virtual_pipeline_model_parallel_size=vpp, sequence_parallel=sequence_parallel and tp > 1,
              params_dtype=dtype, bf16=dtype == torch.bfloat16, fp16=dtype == torch.float16,
              hidden_dropout=dropout, attention_dropout=dropout, add_bias_linear=bias,
              normalization="RMSNorm" if rmsnorm else "LayerNorm", gradient_accumulation_fusion=True,
              pipeline_dtype=dtype, deallocate_pipeline_outputs=True, expert_model_parallel_size=ep,
              num_moe_experts=num_experts, moe_router_topk=moe_topk, moe_grouped_gemm=bool(moe_grouped),
              moe_token_dispatcher_type=(moe_dispatcher or "alltoall") if num_experts else "allgather",
              moe_ffn_hidden_size=moe_ffn, moe_aux_loss_coeff=moe_aux_loss_coeff, fp8=fp8, fp8_recipe=fp8_recipe,
              use_cpu_initialization=not str(device).startswith("cuda"))
    if swiglu:
        kw.update(activation_func=F.silu, gated_linear_unit=True)
    if recompute_layers:
        # full activation recompute of the first N layers of every stage ("block"): trades N/L x one extra forward for memory
        kw.update(recompute_granularity="full", recompute_method="block", recompute_num_layers=int(recompute_layers))
    config = TransformerConfig(**kw)
    if etp is not None:
        kw["expert_tensor_parallel_size"] = etp
    spec = get_gpt_layer_local_spec(num_experts, bool(moe_grouped))
    chunks = []
    n_chunks = vpp or 1
    for c in range(n_chunks):
        if vpp:
            ps.set_virtual_pipeline_model_parallel_rank(c)
        pre, post = ps.is_pipeline_first_stage(), ps.is_pipeline_last_stage()
        m = GPTModel(config, spec, vocab, seq, pre_process=pre, post_process=post,
                     share_embeddings_and_output_weights=True,
                     position_embedding_type="rope" if rope else "learned_absolute")
        for p in m.parameters():
            if not hasattr(p, "tensor_model_parallel"):
                p.tensor_model_parallel = False
        if str(device).startswith("cuda"):
            m = m.to(device)
        if dtype in (torch.bfloat16, torch.float16):
            m = Float16Module(config, m)
        chunks.append(m)
    if vpp:
        ps.set_virtual_pipeline_model_parallel_rank(0)
    ddp_cfg = DistributedDataParallelConfig(grad_reduce_in_fp32=True, overlap_grad_reduce=overlap_grad_reduce,
                                            use_distributed_optimizer=use_distributed_optimizer,
                                            fused_grad_reduce_scatter=fused_grad_rs)
    model = [DistributedDataParallel(config, ddp_cfg, m, disable_bucketing=(i > 0)) for i, m in enumerate(chunks)]
    opt_cfg = OptimizerConfig(lr=lr, weight_decay=0.01, bf16=dtype == torch.bfloat16, fp16=dtype == torch.float16,
                              params_dtype=dtype, use_distributed_optimizer=use_distributed_optimizer,
                              clip_grad=1.0, adam_beta2=0.95,
                              **(dict(use_precision_aware_optimizer=True, exp_avg_dtype=torch.bfloat16,
                                      exp_avg_sq_dtype=torch.bfloat16) if bf16_adam_moments else {}))
    optimizer = get_megatron_optimizer(opt_cfg, model)
    config.grad_scale_func = optimizer.scale_loss
    config.finalize_model_grads_func = finalize_model_grads
    if (fused_tp_comm and sequence_parallel and tp > 1 and torch.cuda.is_available()
            and str(device).startswith("cuda")):
        from .parallel.tp_layers import setup_tp_fused_comm

        f = ffn or 4 * hidden
        setup_tp_fused_comm(seq * micro_batch, hidden, max(f * (2 if swiglu else 1) // tp, hidden))
    if fused_grad_rs and use_distributed_optimizer and dist.is_initialized():
        from .comm.arena import get_arena

        g = ps.get_data_parallel_group(with_context_parallel=True)
        if dist.get_world_size(g) > 1:
            mx = max(buf.max_bucket_numel() for m in model for buf in m.buffers)
            get_arena("dp", g, 2 * mx * 4 + (1 << 20))
    first = ps.is_pipeline_first_stage(ignore_virtual=True)
    last = ps.is_pipeline_last_stage(ignore_virtual=True)
    return BenchState(model, optimizer, config, seq, micro_batch, vocab, first, last)


def synthetic_batch(state: BenchState, device="cuda", seed: int = 0, pinned: bool = False) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(1000 + seed)
    toks = torch.randint(1, state.vocab, (state.micro_batch, state.seq + 1), generator=g, dtype=torch.int64)
    out = {"tokens": toks[:, :-1].contiguous(), "labels": toks[:, 1:].contiguous(),
           "loss_mask": torch.ones(state.micro_batch, state.seq, dtype=torch.float32),
           "position_ids": torch.arange(state.seq, dtype=torch.int64).unsqueeze(0).expand(state.micro_batch, -1).contiguous()}
    if pinned and torch.cuda.is_available():
        return {k: v.pin_memory() for k, v in out.items()}
    return {k: v.to(device) for k, v in out.items()}


def _loss_func(loss_mask, output_tensor):
    losses = output_tensor.float().view(-1)
    lm = loss_mask.view(-1).float()
    tot = lm.sum()
    loss = torch.sum(losses * lm)
    rep = torch.stack([loss.detach(), tot.detach()])
    return loss, tot.to(torch.int), {"lm loss": rep}


def _forward_step(data_iterator, model):
    b = next(data_iterator)
    out = model(b["tokens"], b["position_ids"], None, labels=b["labels"])
    return out, partial(_loss_func, b["loss_mask"])


def train_step(state: BenchState, microbatches: List[Dict[str, torch.Tensor]], sync_loss: bool = True,
               do_optimizer: bool = True) -> Optional[float]:
    """One global step over ``microbatches`` (already on the device, or pinned host tensors which
    are copied with non_blocking H2D as part of the step)."""
    dev = next(state.model[0].parameters()).device

    def gen():
        for mb in microbatches:
            yield {k: (v.to(dev, non_blocking=True) if v.device != dev else v) for k, v in mb.items()}

    for m in state.model:
        m.zero_grad_buffer()
    fb = schedules.get_forward_backward_func()
    iters = [gen() for _ in state.model] if len(state.model) > 1 else gen()
    out = fb(forward_step_func=_forward_step, data_iterator=iters, model=state.model,
             num_microbatches=len(microbatches), seq_length=state.seq, micro_batch_size=state.micro_batch,
             forward_only=False)
    if do_optimizer:
        state.optimizer.step()
    if not sync_loss:
        return None
    if state.post_process and out:
        t = torch.stack([d["lm loss"] for d in out]).sum(0)
        return float((t[0] / t[1]).item())
    return float("nan")
