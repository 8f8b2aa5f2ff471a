"""Epoch-based finetuning driver shared by the classification tasks.

Parity: reference ``tasks/finetune_utils.py`` (``finetune(train_valid_datasets_provider, model_provider,
forward_step, end_of_epoch_callback_provider)``: build loaders, epoch loop on ``train_step``, accuracy
callback at every epoch end, checkpoint at the end) and ``tasks/eval_utils.py`` (``accuracy_func_provider``).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F


def cross_entropy_loss_func(labels: torch.Tensor, output_tensor: torch.Tensor):
    loss = F.cross_entropy(output_tensor.float(), labels)
    return loss, {"lm loss": loss.detach()}


def process_batch(batch: Dict[str, torch.Tensor], device):
    return (batch["text"].long().to(device), batch["types"].long().to(device), batch["label"].long().to(device),
            batch["padding_mask"].long().to(device))


def classification_forward_step(batch, model):
    dev = next(model.parameters()).device
    tokens, types, labels, mask = process_batch(batch, dev)
    logits = model(tokens, mask, tokentype_ids=types)
    return logits, lambda out: cross_entropy_loss_func(labels, out)


@torch.no_grad()
def calculate_correct_answers(model, loader) -> Dict[str, float]:
    was_training = model.training
    model.eval()
    correct = total = 0
    dev = next(model.parameters()).device
    from megatronapp_b200.parallel import state as ps

    mp = ps.model_parallel_is_initialized()
    dp_rank, dp_size = (ps.get_data_parallel_rank(), ps.get_data_parallel_world_size()) if mp else (0, 1)
    for i, batch in enumerate(loader):
        if i % dp_size != dp_rank:               # every replica scores its share; the counts are summed below
            continue
        tokens, types, labels, mask = process_batch(batch, dev)
        pred = model(tokens, mask, tokentype_ids=types).argmax(dim=-1)
        correct += int((pred == labels).sum())
        total += labels.numel()
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        from megatronapp_b200.parallel import state as ps

        g = ps.get_data_parallel_group()
        if g is not None and not isinstance(g, ps._SingleGroup):
            t = torch.tensor([correct, total], device=dev, dtype=torch.float64)
            dist.all_reduce(t, group=g)
            correct, total = int(t[0]), int(t[1])
    model.train(was_training)
    return {"correct": correct, "total": total, "accuracy": correct / max(total, 1)}


def clip_grad_norm_model_parallel(params, max_norm: float) -> float:
    """Global L2 clipping that is consistent across tensor-parallel ranks: sharded parameters contribute their shard,
    replicated ones are counted once (on TP rank 0), and the squared norm is summed over the model-parallel group."""
    import torch.distributed as dist

    from megatronapp_b200.parallel import state as ps
    from megatronapp_b200.parallel.tp_layers import param_is_not_tensor_parallel_duplicate

    grads = [p.grad for p in params if p.grad is not None and
             (not ps.model_parallel_is_initialized() or param_is_not_tensor_parallel_duplicate(p))]
    sq = torch.stack([g.float().pow(2).sum() for g in grads]).sum() if grads else torch.zeros(())
    if ps.model_parallel_is_initialized() and dist.is_initialized():
        g = ps.get_model_parallel_group()
        if g is not None and not isinstance(g, ps._SingleGroup):
            dist.all_reduce(sq, group=g)
    norm = float(sq.sqrt())
    coef = max_norm / (norm + 1e-6)
    if coef < 1.0:
        for p in params:
            if p.grad is not None:
                p.grad.mul_(coef)
    return norm


def finetune(model, train_dataset, valid_dataset, epochs: int = 3, batch_size: int = 16, lr: float = 2e-5,
             weight_decay: float = 0.01, warmup_fraction: float = 0.1, clip_grad: float = 1.0,
             forward_step: Callable = classification_forward_step, end_of_epoch_callback: Optional[Callable] = None,
             log_interval: int = 50, sampler_seed: int = 1234, evaluate: Optional[Callable] = None,
             collate_fn: Optional[Callable] = None, eval_batch_size: Optional[int] = None, keep_last: bool = False):
    """Single-replica (or DDP-sharded through the sampler) finetuning loop; returns per-epoch metrics."""
    import torch.distributed as dist

    from megatronapp_b200.parallel import state as ps

    dp_rank = ps.get_data_parallel_rank() if ps.model_parallel_is_initialized() else 0
    dp_size = ps.get_data_parallel_world_size() if ps.model_parallel_is_initialized() else 1
    sampler = torch.utils.data.distributed.DistributedSampler(train_dataset, dp_size, dp_rank, shuffle=True, seed=sampler_seed)
    loader = torch.utils.data.DataLoader(train_dataset, batch_size=batch_size, sampler=sampler, drop_last=not keep_last, collate_fn=collate_fn)
    vloader = torch.utils.data.DataLoader(valid_dataset, batch_size=eval_batch_size or batch_size, shuffle=False, collate_fn=collate_fn)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=lr, weight_decay=weight_decay)
    total_steps = max(1, epochs * len(loader))
    warm = max(1, int(warmup_fraction * total_steps))
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: min(1.0, (s + 1) / warm) * max(0.0, (total_steps - s) / total_steps))
    history, step = [], 0
    for epoch in range(epochs):
        sampler.set_epoch(epoch)
        model.train()
        running = 0.0
        for batch in loader:
            out, loss_fn = forward_step(batch, model)
            loss, _ = loss_fn(out)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            if dp_size > 1:
                g = ps.get_data_parallel_group()
                for p in params:
                    if p.grad is not None:
                        dist.all_reduce(p.grad, group=g)
                        p.grad.div_(dp_size)
            if clip_grad:
                clip_grad_norm_model_parallel(params, clip_grad)
            opt.step()
            sched.step()
            running += float(loss)
            step += 1
            if log_interval and step % log_interval == 0:
                if not dist.is_initialized() or dist.get_rank() == 0:
                    print(f" epoch {epoch} step {step}/{total_steps} | loss {running / log_interval:.4f} | lr {sched.get_last_lr()[0]:.3e}", flush=True)
                running = 0.0
        metrics = (evaluate or calculate_correct_answers)(model, vloader)
        metrics["epoch"] = epoch
        history.append(metrics)
        if end_of_epoch_callback is not None:
            end_of_epoch_callback(model, epoch, metrics)
    return history
