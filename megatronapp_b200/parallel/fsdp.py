"""Fully-sharded data parallelism (ZeRO-3): parameters, gradients and optimizer state of every
unit (by default each transformer layer, plus one root unit for everything else) live as 1/dp
shards; a unit is all-gathered right before its forward / backward and freed right after, and its
gradient is reduce-scattered as soon as the last parameter of the unit has accumulated.

Parity: reference ``core/distributed/custom_fsdp/`` (``FullyShardedDataParallel``, param/grad
buffer with ``optim_grads_params`` sharding, forward/backward pre-hooks, ``training.py:1144-1148``
selection via ``--use-custom-fsdp``) and ``torch_fully_sharded_data_parallel.py`` (FSDP2 wrapper,
``--use-torch-fsdp2``).  Sized for 180 GB HBM: units are whole layers, so the all-gather unit for
a 6.7B model is ~400 MB bf16 -- large enough to run NVSwitch at full rate.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional

import contextlib
import math

import torch
import torch.distributed as dist
import torch.nn as nn

from . import state as ps


def _all_gather_flat(out: torch.Tensor, shard: torch.Tensor, group):
    if group is None or dist.get_world_size(group) == 1:
        out.copy_(shard)
        return
    dist.all_gather_into_tensor(out, shard, group=group) if dist.get_backend(group) != "gloo" else \
        dist.all_gather(list(out.chunk(dist.get_world_size(group))), shard, group=group)


def _reduce_scatter_flat(out: torch.Tensor, full: torch.Tensor, group):
    if group is None or dist.get_world_size(group) == 1:
        out.copy_(full)
        return
    if dist.get_backend(group) != "gloo":
        dist.reduce_scatter_tensor(out, full, group=group)
    else:
        dist.all_reduce(full, group=group)
        out.copy_(full.chunk(dist.get_world_size(group))[dist.get_rank(group)])


class FSDPUnit:
    """One shard-group: a flat padded buffer holding the unit's parameters back to back."""

    def __init__(self, name: str, params: List[nn.Parameter], group, grad_dtype=torch.float32):
        self.name, self.params, self.group = name, params, group
        self.world = dist.get_world_size(group) if group is not None else 1
        self.rank = dist.get_rank(group) if group is not None else 0
        self.dtype, self.device = params[0].dtype, params[0].device
        n = sum(p.numel() for p in params)
        self.numel = (n + self.world * 8 - 1) // (self.world * 8) * (self.world * 8)
        self.shard_numel = self.numel // self.world
        flat = torch.zeros(self.numel, dtype=self.dtype, device=self.device)
        off = 0
        self.offsets = []
        for p in params:
            flat[off:off + p.numel()].copy_(p.data.reshape(-1))
            self.offsets.append(off)
            off += p.numel()
        lo = self.rank * self.shard_numel
        self.shard = flat[lo:lo + self.shard_numel].clone()                  # model-dtype shard (all-gather source)
        self.main_shard = nn.Parameter(self.shard.float().clone())           # fp32 master, what the optimizer sees
        self.main_shard.grad = None
        self.flat = flat
        self._point_params()
        self.sharded = False
        self.pending = 0
        self.grad_dtype = grad_dtype
        self.reshard()

    def _point_params(self):
        for p, off in zip(self.params, self.offsets):
            p.data = self.flat[off:off + p.numel()].view(p.shape)

    def unshard(self):
        if not self.sharded:
            return
        self.flat.untyped_storage().resize_(self.numel * self.flat.element_size())
        _all_gather_flat(self.flat, self.shard, self.group)
        self.sharded = False

    def reshard(self):
        if self.sharded:
            return
        self.flat.untyped_storage().resize_(0)       # saved-for-backward views see the refill on unshard
        self.sharded = True

    def reduce_grads(self, scale: float):
        g = torch.zeros(self.numel, dtype=self.grad_dtype, device=self.device)
        for p, off in zip(self.params, self.offsets):
            if p.grad is not None:
                g[off:off + p.numel()].copy_(p.grad.reshape(-1))
                p.grad = None
        if scale != 1.0:
            g.mul_(scale)
        out = torch.empty(self.shard_numel, dtype=self.grad_dtype, device=self.device)
        _reduce_scatter_flat(out, g, self.group)
        if self.main_shard.grad is None:
            self.main_shard.grad = out.float()
        else:
            self.main_shard.grad.add_(out.float())

    def copy_main_to_model(self):
        self.shard.copy_(self.main_shard.data)


class FullyShardedDataParallel(nn.Module):
    def __init__(self, config, ddp_config, module: nn.Module, unit_modules: Optional[Iterable[type]] = None,
                 group=None, disable_bucketing: bool = False):
        super().__init__()
        self.module, self.config, self.ddp_config = module, config, ddp_config
        if group is None:
            g = ps.get_data_parallel_group(with_context_parallel=True)
            group = None if isinstance(g, ps._SingleGroup) else g
        self.group = group
        world = dist.get_world_size(group) if group is not None else 1
        self.scale = 1.0 / world
        if unit_modules is None:
            from ..transformer.transformer_layer import TransformerLayer

            unit_modules = (TransformerLayer,)
        unit_modules = tuple(unit_modules)
        self.units: List[FSDPUnit] = []
        claimed = set()
        self._unit_of_module = {}
        for name, m in module.named_modules():
            if isinstance(m, unit_modules):
                ps_ = [p for p in m.parameters() if p.requires_grad and id(p) not in claimed]
                if ps_:
                    claimed.update(id(p) for p in ps_)
                    u = FSDPUnit(name, ps_, group)
                    self.units.append(u)
                    self._unit_of_module[m] = u
        rest = [p for p in module.parameters() if p.requires_grad and id(p) not in claimed]
        self.root_unit = FSDPUnit("root", rest, group) if rest else None
        if self.root_unit is not None:
            self.units.append(self.root_unit)
        for m, u in self._unit_of_module.items():
            m.register_forward_pre_hook(lambda mod, args, u=u: u.unshard())
            m.register_forward_hook(lambda mod, args, out, u=u: u.reshard() if torch.is_grad_enabled() or True else None)
            m.register_full_backward_pre_hook(lambda mod, gout, u=u: self._pre_backward(u))
        for u in self.units:
            for p in u.params:
                p.register_post_accumulate_grad_hook(lambda p_, u=u: self._grad_ready(u))

    # -- hooks -----------------------------------------------------------------------------
    def _pre_backward(self, u: FSDPUnit):
        u.unshard()
        if u.pending == 0:
            u.pending = len(u.params)

    def _grad_ready(self, u: FSDPUnit):
        if u.pending == 0:       # root unit (no module-level pre-backward hook)
            u.pending = len(u.params)
        u.pending -= 1
        if u.pending == 0:
            u.reduce_grads(self.scale)
            if u is not self.root_unit:
                u.reshard()

    def forward(self, *a, **k):
        if self.root_unit is not None:
            self.root_unit.unshard()
        return self.module(*a, **k)

    # -- DDP-compatible surface used by the training loop ---------------------------------------
    def zero_grad_buffer(self):
        for u in self.units:
            u.main_shard.grad = None
            u.pending = 0

    def finish_grad_sync(self):
        for u in self.units:      # params that never received a gradient this step
            if u.pending not in (0, len(u.params)):
                u.reduce_grads(self.scale)
                u.pending = 0
        if self.root_unit is not None:
            self.root_unit.reshard()

    start_grad_sync = finish_grad_sync

    def no_sync(self):
        import contextlib

        return contextlib.nullcontext()

    def shard_parameters(self) -> List[nn.Parameter]:
        return [u.main_shard for u in self.units]

    def post_optimizer_step(self):
        for u in self.units:
            u.copy_main_to_model()

    def broadcast_params(self):
        pass

    @contextlib.contextmanager
    def summon_full_params(self, writeback: bool = False):
        """All-gather every unit so that the wrapped module's parameters are whole (checkpoint save / load, export);
        with ``writeback`` the possibly modified full parameters are folded back into the local shards on exit."""
        for u in self.units:
            u.unshard()
        try:
            yield self.module
        finally:
            for u in self.units:
                if writeback:
                    lo = u.rank * u.shard_numel
                    u.shard.copy_(u.flat[lo:lo + u.shard_numel])
                    u.main_shard.data.copy_(u.shard.float())
                u.reshard()

    def state_dict_for_save_checkpoint(self, prefix="", keep_vars=False):
        return self.state_dict()

    def state_dict(self, *a, **k):
        for u in self.units:
            u.unshard()
        sd = {k_: v.clone() for k_, v in self.module.state_dict(*a, **k).items()}
        for u in self.units:
            u.reshard()
        return sd

    def load_state_dict(self, sd, strict=True):
        for u in self.units:
            u.unshard()
        r = self.module.load_state_dict(sd, strict=strict)
        for u in self.units:
            lo = u.rank * u.shard_numel
            u.shard.copy_(u.flat[lo:lo + u.shard_numel])
            u.main_shard.data.copy_(u.shard.float())
            u.reshard()
        return r


class TorchFullyShardedDataParallel(nn.Module):
    """Thin wrapper over torch FSDP2 ``fully_shard`` (reference ``torch_fully_sharded_data_parallel.py``)."""

    def __init__(self, config, ddp_config, module: nn.Module, sub_modules_to_wrap: Optional[Iterable[type]] = None,
                 group=None):
        super().__init__()
        from torch.distributed.fsdp import fully_shard
        from torch.distributed.device_mesh import DeviceMesh

        if sub_modules_to_wrap is None:
            from ..transformer.transformer_layer import TransformerLayer

            sub_modules_to_wrap = (TransformerLayer,)
        g = group or ps.get_data_parallel_group(with_context_parallel=True)
        mesh = DeviceMesh.from_group(g, "cuda" if torch.cuda.is_available() else "cpu")
        for m in module.modules():
            if isinstance(m, tuple(sub_modules_to_wrap)):
                fully_shard(m, mesh=mesh)
        fully_shard(module, mesh=mesh)
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)

    def zero_grad_buffer(self):
        self.module.zero_grad(set_to_none=True)

    def finish_grad_sync(self):
        pass

    start_grad_sync = finish_grad_sync
    broadcast_params = finish_grad_sync


class FSDPOptimizer:
    """Adam(W) on the fp32 master shards of ``FullyShardedDataParallel`` chunks (fused sm_100a Adam
    kernel on CUDA), global-norm clipping across dp x model-parallel, same surface as
    ``optim.MegatronOptimizer`` so the training loop does not care which wrapper is in use."""

    def __init__(self, config, model_chunks):
        import math

        from .. import ops

        self.config, self.chunks, self.ops, self.math = config, list(model_chunks), ops, math
        self.units = [u for c in self.chunks for u in c.units]
        self.state = [{"m": torch.zeros_like(u.main_shard.data), "v": torch.zeros_like(u.main_shard.data)} for u in self.units]
        self.param_groups = [{"name": "all", "lr": config.lr, "weight_decay": config.weight_decay, "wd_mult": 1.0,
                              "lr_mult": 1.0, "params": [u.main_shard for u in self.units]}]
        self.step_count = 0
        self.sharded = True

    def zero_grad(self, set_to_none=True):
        for c in self.chunks:
            c.zero_grad_buffer()

    def get_loss_scale(self):
        return torch.ones(1)

    def scale_loss(self, loss):
        return loss

    def get_grad_norm(self) -> float:
        grads = [u.main_shard.grad for u in self.units if u.main_shard.grad is not None]
        sq = self.ops.l2norm_sq(grads) if grads else torch.zeros(1)
        g = self.chunks[0].group
        if g is not None:
            dist.all_reduce(sq, group=g)
        mp = ps.get_model_parallel_group()
        if mp is not None and not isinstance(mp, ps._SingleGroup):
            dist.all_reduce(sq, group=mp)
        return float(sq.sqrt())

    def step(self):
        for c in self.chunks:
            c.finish_grad_sync()
        self.step_count += 1
        cfg = self.config
        norm = self.get_grad_norm()
        clip = 1.0
        if cfg.clip_grad and cfg.clip_grad > 0 and norm > cfg.clip_grad:
            clip = cfg.clip_grad / (norm + 1e-6)
        g = self.param_groups[0]
        for u, st in zip(self.units, self.state):
            if u.main_shard.grad is None:
                continue
            self.ops.adam_step(u.main_shard.data, st["m"], st["v"], u.main_shard.grad, u.shard, g["lr"], cfg.adam_beta1,
                               cfg.adam_beta2, cfg.adam_eps, g["weight_decay"], clip, self.step_count)
        return True, norm, None

    def reload_model_params(self):
        for u in self.units:
            u.main_shard.data.copy_(u.shard.float())

    def state_dict(self):
        return {"step": self.step_count, "state": [{k: v.cpu() for k, v in s.items()} for s in self.state],
                "main": [u.main_shard.data.cpu() for u in self.units]}

    def load_state_dict(self, sd):
        self.step_count = sd["step"]
        for s, t in zip(self.state, sd["state"]):
            for k in s:
                s[k].copy_(t[k])
        for u, m in zip(self.units, sd["main"]):
            u.main_shard.data.copy_(m)
            u.copy_main_to_model()


class TorchFSDPOptimizer:
    """AdamW over the DTensor parameters of torch-FSDP2-wrapped chunks, behind the interface the training loop expects
    (``step() -> (ok, grad_norm, num_zeros)``, ``param_groups`` with lr / weight decay set by the scheduler,
    ``scale_loss``, ``state_dict``).  torch does the sharded bookkeeping; this only adapts the surface."""

    sharded = True

    def __init__(self, config, model_chunks):
        self.config, self.model_chunks = config, model_chunks
        decay, no_decay = [], []
        for m in model_chunks:
            for n, p in m.named_parameters():
                if p.requires_grad:
                    (no_decay if (p.dim() <= 1 or n.endswith(".bias")) else decay).append(p)
        self.opt = torch.optim.AdamW([{"params": decay, "weight_decay": config.weight_decay, "wd_mult": 1.0, "lr_mult": 1.0},
                                      {"params": no_decay, "weight_decay": 0.0, "wd_mult": 0.0, "lr_mult": 1.0}],
                                     lr=config.lr or 0.0, betas=(config.adam_beta1, config.adam_beta2), eps=config.adam_eps)
        self.param_groups = self.opt.param_groups
        self.grad_scaler = None

    def zero_grad(self, set_to_none=True):
        self.opt.zero_grad(set_to_none=set_to_none)

    def get_loss_scale(self):
        return torch.tensor([1.0])

    def scale_loss(self, loss):
        return loss

    def _params(self):
        return [p for g in self.opt.param_groups for p in g["params"] if p.grad is not None]

    @torch.no_grad()
    def step(self):
        params = self._params()
        norm = None
        if self.config.clip_grad > 0.0 and params:
            total = torch.nn.utils.clip_grad_norm_(params, self.config.clip_grad)
            total = total.full_tensor() if hasattr(total, "full_tensor") else total
            norm = float(total)
            if not math.isfinite(norm):
                return False, norm, None
        self.opt.step()
        return True, norm, None

    def reload_model_params(self):
        pass

    def state_dict(self):
        return self.opt.state_dict()

    def load_state_dict(self, sd):
        self.opt.load_state_dict(sd)

    def sharded_state_dict(self, *a, **k):
        return self.state_dict()
