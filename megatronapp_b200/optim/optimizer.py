"""Mixed-precision Adam(W) on the contiguous DDP buffers, replicated or ZeRO-1 sharded.

Parity: reference ``core/optimizer/__init__.py`` (``get_megatron_optimizer``),
``optimizer.py`` (``Float16OptimizerWithFloat16Params`` :471, ``FP32Optimizer``,
``ChainedOptimizer``), ``distrib_optimizer.py`` (ZeRO-1: each DP rank owns 1/dp of every
bucket: grad reduce-scatter, local Adam on the fp32 main shard, param all-gather),
``clip_grads.py`` (global L2 norm with TP-duplicate / shared-param exclusion + MP all-reduce),
``grad_scaler.py`` (constant / dynamic loss scaling), ``optimizer_config.py``.

B200-first design: because every param is a view into one contiguous ``param_data`` buffer and
every ``main_grad`` a view into ``grad_data``, the whole update is a handful of launches of ONE
fused kernel (``adam_step``: fp32 master + moments update, decoupled weight decay, clip /
unscale folded into ``grad_scale``, bf16 param write-back) over merged ``(weight-decay,
lr-mult)`` ranges -- no multi-tensor-apply lists, no separate main->model copy pass.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from ..parallel import state as ps
from ..parallel.ddp import DistributedDataParallel, ParamAndGradBuffer, _is_real_group
from .. import ops
from ..scan.tracer import scope as _scope


@dataclass
class OptimizerConfig:
    optimizer: str = "adam"
    lr: Optional[float] = None
    min_lr: Optional[float] = None
    decoupled_lr: Optional[float] = None
    decoupled_min_lr: Optional[float] = None
    weight_decay: float = 0.01
    fp16: bool = False
    bf16: bool = False
    params_dtype: torch.dtype = torch.float32
    use_precision_aware_optimizer: bool = False
    main_grads_dtype: torch.dtype = torch.float32
    main_params_dtype: torch.dtype = torch.float32
    exp_avg_dtype: torch.dtype = torch.float32
    exp_avg_sq_dtype: torch.dtype = torch.float32
    loss_scale: Optional[float] = None
    initial_loss_scale: float = 2 ** 32
    min_loss_scale: float = 1.0
    loss_scale_window: float = 1000
    hysteresis: int = 2
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_eps: float = 1e-8
    sgd_momentum: float = 0.9
    use_distributed_optimizer: bool = False
    overlap_param_gather_with_optimizer_step: bool = False
    optimizer_cpu_offload: bool = False
    optimizer_offload_fraction: float = 0.0
    clip_grad: float = 1.0
    log_num_zeros_in_grad: bool = False
    barrier_with_L1_time: bool = False
    timers: Optional[Callable] = None
    config_logger_dir: str = ""
    # B200: param all-gather fused with the fp32->bf16 cast over the DP arena
    fused_param_all_gather: bool = False


class ConstantGradScaler:
    def __init__(self, scale: float):
        self._scale = torch.tensor([scale], dtype=torch.float32,
                                   device="cuda" if torch.cuda.is_available() else "cpu")

    @property
    def scale(self):
        return self._scale

    @property
    def inv_scale(self):
        return self._scale.double().reciprocal().float()

    def update(self, found_inf: bool):
        pass

    def state_dict(self):
        return {}

    def load_state_dict(self, sd):
        pass


class DynamicGradScaler(ConstantGradScaler):
    def __init__(self, initial_scale, min_scale, growth_factor, backoff_factor, growth_interval, hysteresis):
        super().__init__(initial_scale)
        self.min_scale = min_scale
        self.growth_factor, self.backoff_factor = growth_factor, backoff_factor
        self.growth_interval, self.hysteresis = growth_interval, hysteresis
        self._growth_tracker = 0
        self._hysteresis_tracker = hysteresis

    def update(self, found_inf: bool):
        if found_inf:
            self._growth_tracker = 0
            self._hysteresis_tracker -= 1
            if self._hysteresis_tracker <= 0:
                self._scale = torch.clamp(self._scale * self.backoff_factor, min=self.min_scale)
        else:
            self._growth_tracker += 1
            if self._growth_tracker == self.growth_interval:
                self._growth_tracker = 0
                self._hysteresis_tracker = self.hysteresis
                self._scale = self._scale * self.growth_factor

    def state_dict(self):
        return {"scale": self._scale, "growth_tracker": self._growth_tracker,
                "hysteresis_tracker": self._hysteresis_tracker}

    def load_state_dict(self, sd):
        self._scale = sd["scale"].to(self._scale.device)
        self._growth_tracker = sd["growth_tracker"]
        self._hysteresis_tracker = sd["hysteresis_tracker"]


def default_no_weight_decay_cond(name: str, param) -> bool:
    return name.endswith(".bias") or param.dim() == 1


@dataclass
class _Range:
    start: int
    end: int
    wd_mult: float
    lr_mult: float


class _BufferState:
    """Optimizer state for one ParamAndGradBuffer (full, or this rank's shards under ZeRO-1)."""

    def __init__(self, buf: ParamAndGradBuffer, ranges: List[_Range], sharded: bool, offload_fraction: float = 0.0,
                 moment_dtypes=(torch.float32, torch.float32)):
        self.buf, self.sharded = buf, sharded
        dev = buf.param_data.device
        if sharded:
            self.segments = [buf.shard_range(b) for b in buf.buckets]
        else:
            self.segments = [(0, buf.numel)]
        self.seg_offsets = []
        off = 0
        for s, e in self.segments:
            self.seg_offsets.append(off)
            off += e - s
        self.local_numel = off
        self.master = torch.empty(off, dtype=torch.float32, device=dev)
        for (s, e), o in zip(self.segments, self.seg_offsets):
            self.master[o:o + e - s].copy_(buf.param_data[s:e])
        # precision-aware optimizer: the moments may be stored in 16 bit (half the optimizer-state memory); the step
        # up-casts a slice, runs the fp32 Adam arithmetic and stores it back
        self.exp_avg = torch.zeros_like(self.master, dtype=moment_dtypes[0])
        self.exp_avg_sq = torch.zeros_like(self.master, dtype=moment_dtypes[1])
        # --optimizer-cpu-offload: the tail [cpu_from, local_numel) of master / moments lives in (pinned) host memory and
        # is stepped on the CPU (reference optimizer/cpu_offloading/hybrid_optimizer.py: HybridDeviceOptimizer)
        self.cpu_from = off
        self.master_cpu = self.exp_avg_cpu = self.exp_avg_sq_cpu = None
        if offload_fraction > 0.0 and off > 0:
            keep = int(off * (1.0 - min(1.0, offload_fraction))) // 128 * 128
            pin = dev.type == "cuda"
            host = lambda t: (t.cpu().pin_memory() if pin else t.cpu().clone())  # noqa: E731
            self.master_cpu = host(self.master[keep:])
            self.exp_avg_cpu, self.exp_avg_sq_cpu = host(self.exp_avg[keep:]), host(self.exp_avg_sq[keep:])
            self.master, self.exp_avg, self.exp_avg_sq = (t[:keep].clone() for t in (self.master, self.exp_avg, self.exp_avg_sq))
            self.cpu_from = keep
        # intersect (wd, lr) ranges with owned segments -> launch list of (buf_start, buf_end, local_off, wd, lr)
        self.launches: List[Tuple[int, int, int, float, float]] = []
        for (s, e), o in zip(self.segments, self.seg_offsets):
            for r in ranges:
                a, b = max(s, r.start), min(e, r.end)
                if a < b:
                    lo = o + (a - s)
                    cut = min(max(self.cpu_from - lo, 0), b - a)        # elements of this launch that stay on the device
                    if cut > 0:
                        self.launches.append((a, a + cut, lo, r.wd_mult, r.lr_mult))
                    if cut < b - a:
                        self.launches.append((a + cut, b, lo + cut, r.wd_mult, r.lr_mult))

    def full(self, name: str) -> torch.Tensor:
        """Device part + offloaded part of ``master`` / ``exp_avg`` / ``exp_avg_sq`` as one CPU-or-device tensor."""
        dev_part, cpu_part = getattr(self, name), getattr(self, name + "_cpu")
        return dev_part if cpu_part is None else torch.cat([dev_part.cpu(), cpu_part])

    def load_full(self, name: str, value: torch.Tensor):
        dev_part, cpu_part = getattr(self, name), getattr(self, name + "_cpu")
        dev_part.copy_(value[: dev_part.numel()])
        if cpu_part is not None:
            cpu_part.copy_(value[dev_part.numel():])


def _adam(master, m, v, grad, p16, lr, cfg, wd, gs, step):
    """One fused-Adam launch on a slice; 16-bit moments are up-cast around the fp32 kernel and stored back."""
    low = m.dtype != torch.float32 or v.dtype != torch.float32
    m32, v32 = (m.float(), v.float()) if low else (m, v)
    ops.adam_step(master, m32, v32, grad, p16, lr, cfg.adam_beta1, cfg.adam_beta2, cfg.adam_eps, wd, gs, step)
    if low:
        m.copy_(m32)
        v.copy_(v32)


class MegatronOptimizer:
    """Adam(W) over the DDP buffers of a list of model chunks."""

    def __init__(self, config: OptimizerConfig, model_chunks: List[DistributedDataParallel],
                 no_weight_decay_cond: Optional[Callable] = None, scale_lr_cond: Optional[Callable] = None,
                 lr_mult: float = 1.0):
        self.config = config
        self.model_chunks = model_chunks
        self.step_count = 0
        self.sharded = config.use_distributed_optimizer
        nwd = no_weight_decay_cond or default_no_weight_decay_cond
        self.param_groups = [
            {"name": "wd", "wd_mult": 1.0, "lr_mult": 1.0, "lr": config.lr, "weight_decay": config.weight_decay,
             "is_decoupled_lr": False},
            {"name": "no_wd", "wd_mult": 0.0, "lr_mult": 1.0, "lr": config.lr, "weight_decay": 0.0,
             "is_decoupled_lr": False},
        ]
        self.states: List[_BufferState] = []
        self._excluded_from_norm: List[torch.nn.Parameter] = []
        for chunk in model_chunks:
            names = {p: n for n, p in chunk.module.named_parameters()}
            for buf in chunk.buffers + chunk.expert_parallel_buffers:
                is_expert_buf = any(buf is b for b in chunk.expert_parallel_buffers)
                raw = []
                for p in buf.params:
                    s, e, _ = buf.param_index[p]
                    n = names.get(p, "")
                    wd = 0.0 if nwd(n, p) else 1.0
                    lm = lr_mult if (scale_lr_cond is not None and scale_lr_cond(n, p)) else 1.0
                    raw.append(_Range(s, e, wd, lm))
                    is_dup = (not getattr(p, "tensor_model_parallel", False)) and ps.get_tensor_model_parallel_rank() != 0
                    if is_dup or getattr(p, "shared", False):
                        self._excluded_from_norm.append(p)
                raw.sort(key=lambda r: r.start)
                merged: List[_Range] = []
                for r in raw:
                    if merged and merged[-1].wd_mult == r.wd_mult and merged[-1].lr_mult == r.lr_mult:
                        merged[-1].end = r.end  # swallow alignment padding in between
                    else:
                        merged.append(_Range(r.start, r.end, r.wd_mult, r.lr_mult))
                self.states.append(_BufferState(buf, merged, self.sharded and _is_real_group(buf.group),
                                                config.optimizer_offload_fraction if config.optimizer_cpu_offload else 0.0,
                                                (config.exp_avg_dtype, config.exp_avg_sq_dtype)
                                                if config.use_precision_aware_optimizer else (torch.float32, torch.float32)))
                self.states[-1].is_expert = is_expert_buf
        if config.fp16:
            if config.loss_scale:
                self.grad_scaler = ConstantGradScaler(config.loss_scale)
            else:
                self.grad_scaler = DynamicGradScaler(config.initial_loss_scale, config.min_loss_scale, 2.0, 0.5,
                                                     config.loss_scale_window, config.hysteresis)
        else:
            self.grad_scaler = ConstantGradScaler(config.loss_scale) if config.loss_scale else None
        dev = self.states[0].master.device if self.states else "cpu"
        self._found_inf = torch.zeros(1, dtype=torch.float32, device=dev)
        self._ag_epoch = 0

    # -- loss scaling ---------------------------------------------------------------
    def get_loss_scale(self):
        if self.grad_scaler is None:
            return torch.tensor([1.0], dtype=torch.float32, device=self._found_inf.device)
        return self.grad_scaler.scale

    def scale_loss(self, loss):
        return self.get_loss_scale() * loss

    def zero_grad(self, set_to_none=True):
        for chunk in self.model_chunks:
            for p in chunk.module.parameters():
                p.grad = None

    # -- gradient statistics ------------------------------------------------------------
    def _owned_grads(self):
        for st in self.states:
            for s, e in st.segments:
                yield st.buf.grad_data[s:e]

    def _norm_sq(self, states) -> torch.Tensor:
        """Sum of squares of the grads owned by ``states`` minus the TP-duplicated / shared copies."""
        dev = self._found_inf.device
        total = torch.zeros(1, dtype=torch.float32, device=dev)
        grads = [st.buf.grad_data[s:e] for st in states for s, e in st.segments]
        if grads:
            ops.l2norm_sq(grads, total)
        excl_params = [p for st in states for p in st.buf.params if any(p is q for q in self._excluded_from_norm)]
        if excl_params:
            excl = torch.zeros(1, dtype=torch.float32, device=dev)
            if self.sharded:
                # excluded params may straddle shards: subtract only the owned part
                parts = []
                for st in states:
                    for p in st.buf.params:
                        if any(p is q for q in excl_params):
                            s, e, _ = st.buf.param_index[p]
                            for a, b in st.segments:
                                lo, hi = max(a, s), min(b, e)
                                if lo < hi:
                                    parts.append(st.buf.grad_data[lo:hi])
                ops.l2norm_sq(parts, excl) if parts else None
            else:
                ops.l2norm_sq([p.main_grad.reshape(-1) for p in excl_params], excl)
            total = total - excl
        if self.sharded:
            for st in states:
                if _is_real_group(st.buf.group):
                    dist.all_reduce(total, group=st.buf.group)
                    break
        return total

    def _model_parallel_groups(self):
        """(dense group, expert group): dense grads are distinct across tp x pp; expert grads are distinct across
        tp x ep x pp (every EP rank holds different experts) -- reference ``clip_grads`` / ``get_grad_stats_parallel_group``."""
        if not ps.model_parallel_is_initialized():
            return None, None
        mp = ps.get_model_parallel_group(check_initialized=False)
        try:
            emp = ps.get_state().group("tp_ep_pp")
        except Exception:
            emp = mp
        return mp, emp

    def get_grad_norm(self) -> float:
        dense = [st for st in self.states if not st.is_expert]
        expert = [st for st in self.states if st.is_expert]
        mp, emp = self._model_parallel_groups()
        total = self._norm_sq(dense)
        if _is_real_group(mp):
            dist.all_reduce(total, group=mp)
        if expert:
            etotal = self._norm_sq(expert)
            if _is_real_group(emp):
                dist.all_reduce(etotal, group=emp)
            elif _is_real_group(mp):
                dist.all_reduce(etotal, group=mp)
            total = total + etotal
        return float(total.clamp_min(0).sqrt().item())

    def count_zeros(self) -> float:
        n = torch.zeros(1, dtype=torch.float32, device=self._found_inf.device)
        for g in self._owned_grads():
            n += (g == 0).sum()
        return float(n.item())

    def _check_found_inf(self) -> bool:
        self._found_inf.zero_()
        for g in self._owned_grads():
            if g.is_cuda:
                ops.ext().check_finite(g, self._found_inf)
            elif not torch.isfinite(g).all():
                self._found_inf.fill_(1.0)
        mp, emp = self._model_parallel_groups()
        if _is_real_group(mp):
            dist.all_reduce(self._found_inf, op=dist.ReduceOp.MAX, group=mp)
        if any(st.is_expert for st in self.states) and _is_real_group(emp) and emp is not mp:
            # an inf that only one EP rank's experts produced must stop the step on every rank
            dist.all_reduce(self._found_inf, op=dist.ReduceOp.MAX, group=emp)
        if self.sharded:
            for st in self.states:
                if _is_real_group(st.buf.group):
                    dist.all_reduce(self._found_inf, op=dist.ReduceOp.MAX, group=st.buf.group)
                    break
        return self._found_inf.item() > 0

    # -- the step ---------------------------------------------------------------------
    @torch.no_grad()
    def step(self):
        cfg = self.config
        inv_scale = 1.0
        if self.grad_scaler is not None:
            found_inf = self._check_found_inf()
            self.grad_scaler.update(found_inf)
            if found_inf:
                return False, None, None
            inv_scale = 1.0 / float(self.grad_scaler.scale.item())
        grad_norm = None
        clip = 1.0
        if cfg.clip_grad > 0.0:
            grad_norm = self.get_grad_norm() * inv_scale
            if not math.isfinite(grad_norm):
                return False, grad_norm, None
            coeff = cfg.clip_grad / (grad_norm + 1.0e-6)
            if coeff < 1.0:
                clip = coeff
        num_zeros = self.count_zeros() if cfg.log_num_zeros_in_grad else None
        self.step_count += 1
        gs = inv_scale * clip
        for st in self.states:
            buf = st.buf
            for (a, b, lo, wd_mult, lr_mult) in st.launches:
                n = b - a
                group = self.param_groups[0 if wd_mult > 0 else 1]
                lr = group["lr"] * lr_mult
                wd = group["weight_decay"] if wd_mult > 0 else 0.0
                grad = buf.grad_data[a:b]
                if grad.dtype != torch.float32:
                    grad = grad.float()
                if lo >= st.cpu_from:                      # offloaded slice: D2H grad, CPU Adam, H2D parameter
                    c = lo - st.cpu_from
                    _adam(st.master_cpu[c:c + n], st.exp_avg_cpu[c:c + n], st.exp_avg_sq_cpu[c:c + n],
                          grad.to("cpu", non_blocking=False), None, lr, cfg, wd, gs, self.step_count)
                    buf.param_data[a:b].copy_(st.master_cpu[c:c + n], non_blocking=True)
                    continue
                p16 = buf.param_data[a:b] if buf.param_dtype != torch.float32 else None
                _adam(st.master[lo:lo + n], st.exp_avg[lo:lo + n], st.exp_avg_sq[lo:lo + n], grad, p16, lr, cfg, wd, gs,
                      self.step_count)
                if p16 is None:
                    buf.param_data[a:b].copy_(st.master[lo:lo + n])
        if self.sharded:
            self._all_gather_params()
        return True, grad_norm, num_zeros

    def _all_gather_params(self):
        with _scope("param-all-gather"):
            for st in self.states:
                buf = st.buf
                if not st.sharded:
                    continue
                for b in buf.buckets:
                    s, e = buf.shard_range(b)
                    dist.all_gather_into_tensor(buf.param_data[b.start:b.end], buf.param_data[s:e], group=buf.group)

    # -- (de)serialisation ----------------------------------------------------------------
    def reload_model_params(self):
        for st in self.states:
            full = torch.empty(st.local_numel, dtype=torch.float32, device=st.buf.param_data.device)
            for (s, e), o in zip(st.segments, st.seg_offsets):
                full[o:o + e - s].copy_(st.buf.param_data[s:e])
            st.load_full("master", full)

    def state_dict(self):
        sd = {"step": self.step_count,
              "param_groups": [{k: v for k, v in g.items()} for g in self.param_groups],
              "buffers": [{"master": st.full("master"), "exp_avg": st.full("exp_avg"), "exp_avg_sq": st.full("exp_avg_sq"),
                           "segments": st.segments} for st in self.states]}
        if self.grad_scaler is not None:
            sd["grad_scaler"] = self.grad_scaler.state_dict()
        return sd

    def load_state_dict(self, sd):
        self.step_count = sd["step"]
        for g, sg in zip(self.param_groups, sd.get("param_groups", [])):
            g.update({k: v for k, v in sg.items() if k in ("lr", "weight_decay")})
        for st, s in zip(self.states, sd["buffers"]):
            st.load_full("master", s["master"])
            st.load_full("exp_avg", s["exp_avg"])
            st.load_full("exp_avg_sq", s["exp_avg_sq"])
        if self.grad_scaler is not None and "grad_scaler" in sd:
            self.grad_scaler.load_state_dict(sd["grad_scaler"])

    def sharded_state_dict(self, model_sharded_state_dict=None, is_loading=False, **kw):
        """Optimizer state in MODEL SPACE, the reference's layout (``dist_checkpointing/optimizer.py:75-140``,
        ``distrib_optimizer.sharded_param_state_fs_model_space``): for every model ShardedTensor ``K`` the fp32 main
        parameter and both Adam moments are ShardedTensors ``optimizer.state.{param|fp32_param,exp_avg,exp_avg_sq}.K`` with
        the parameter's own sharding; under the distributed optimizer each rank contributes the run of the flattened
        local parameter it owns (``flattened_range``).  Such a checkpoint loads under any TP / PP / DP layout."""
        from dataclasses import replace

        from ..ckpt.sharded import ShardedTensor

        if model_sharded_state_dict is None:
            return self.state_dict()
        if any(st.master_cpu is not None for st in self.states):
            raise NotImplementedError("sharded optimizer state with --optimizer-cpu-offload: use --ckpt-format torch")
        owners = []                      # (ptr_lo, ptr_hi, state, param, buffer start)
        for st in self.states:
            for p in st.buf.params:
                s, e, _ = st.buf.param_index[p]
                owners.append((p.data_ptr(), p.data_ptr() + p.numel() * p.element_size(), st, p, s))
        owners.sort(key=lambda o: o[0])
        import bisect

        los = [o[0] for o in owners]
        out = {}
        main_name = "param" if self.sharded else "fp32_param"
        for mkey, mst in model_sharded_state_dict.items():
            if not isinstance(mst, ShardedTensor) or mst.data is None or not torch.is_tensor(mst.data):
                continue
            ptr = mst.data.data_ptr()
            i = bisect.bisect_right(los, ptr) - 1
            if i < 0 or ptr >= owners[i][1] or not mst.data.is_contiguous():
                continue                  # buffers / frozen tensors: no optimizer state
            _, _, st, p, s = owners[i]
            off = (ptr - p.data_ptr()) // p.element_size()
            a, b = s + off, s + off + mst.data.numel()
            for name, attr in ((main_name, "master"), ("exp_avg", "exp_avg"), ("exp_avg_sq", "exp_avg_sq")):
                flat = getattr(st, attr)
                key = f"optimizer.state.{name}.{mst.key}"
                if not st.sharded:
                    out[f"optimizer.state.{name}.{mkey}"] = replace(mst, key=key, data=flat[a:b].view(mst.data.shape),
                                                                     dtype=flat.dtype)
                    continue
                for (ss, se), o in zip(st.segments, st.seg_offsets):
                    lo, hi = max(a, ss), min(b, se)
                    if lo < hi:
                        rid = tuple(mst.replica_id) if isinstance(mst.replica_id, (tuple, list)) else (mst.replica_id,)
                        out[f"optimizer.state.{name}.{mkey}"] = replace(
                            mst, key=key, data=flat[o + lo - ss:o + hi - ss], dtype=flat.dtype,
                            flattened_range=slice(lo - a, hi - a), replica_id=rid[:-1] + (0,), replica_count=1)
        return out

    def supports_model_space_state(self) -> bool:
        """False when part of the state lives in pinned host memory (--optimizer-cpu-offload): such runs checkpoint the
        optimizer per rank (same-layout restore), like the ZeRO-3 wrappers."""
        return not any(st.master_cpu is not None for st in self.states)

    def sharded_common_state(self) -> dict:
        """The non-tensor part of the optimizer state (goes to ``common.pt``)."""
        sd = {"step": self.step_count, "param_groups": [{k: v for k, v in g.items()} for g in self.param_groups]}
        if self.grad_scaler is not None:
            sd["grad_scaler"] = self.grad_scaler.state_dict()
        return sd

    def load_sharded_common_state(self, sd: dict) -> None:
        self.step_count = sd.get("step", 0)
        for g, sg in zip(self.param_groups, sd.get("param_groups", [])):
            g.update({k: v for k, v in sg.items() if k in ("lr", "weight_decay")})
        if self.grad_scaler is not None and "grad_scaler" in sd:
            self.grad_scaler.load_state_dict(sd["grad_scaler"])


Float16OptimizerWithFloat16Params = MegatronOptimizer
DistributedOptimizer = MegatronOptimizer
FP32Optimizer = MegatronOptimizer


class ChainedOptimizer:
    """Several optimizers stepped together (dense + expert parameter groups)."""

    def __init__(self, chained: List[MegatronOptimizer]):
        self.chained_optimizers = chained

    @property
    def param_groups(self):
        out = []
        for o in self.chained_optimizers:
            out += o.param_groups
        return out

    def zero_grad(self, set_to_none=True):
        for o in self.chained_optimizers:
            o.zero_grad(set_to_none)

    def get_loss_scale(self):
        return self.chained_optimizers[0].get_loss_scale()

    def scale_loss(self, loss):
        return self.chained_optimizers[0].scale_loss(loss)

    def step(self):
        ok, norms, zeros = True, [], 0
        for o in self.chained_optimizers:
            s, n, z = o.step()
            ok = ok and s
            if n is not None:
                norms.append(n)
            zeros += z or 0
        return ok, (math.sqrt(sum(x * x for x in norms)) if norms else None), zeros

    def reload_model_params(self):
        for o in self.chained_optimizers:
            o.reload_model_params()

    def state_dict(self):
        return [o.state_dict() for o in self.chained_optimizers]

    def load_state_dict(self, sd):
        for o, s in zip(self.chained_optimizers, sd):
            o.load_state_dict(s)

    def sharded_state_dict(self, model_sharded_state_dict=None, is_loading=False, **kw):
        if model_sharded_state_dict is None:
            return self.state_dict()
        out = {}
        for o in self.chained_optimizers:
            out.update(o.sharded_state_dict(model_sharded_state_dict, is_loading, **kw))
        return out

    def supports_model_space_state(self) -> bool:
        return all(getattr(o, "supports_model_space_state", lambda: True)() for o in self.chained_optimizers)

    def sharded_common_state(self):
        return [o.sharded_common_state() for o in self.chained_optimizers]

    def load_sharded_common_state(self, sd):
        for o, s in zip(self.chained_optimizers, sd):
            o.load_sharded_common_state(s)


def get_megatron_optimizer(config: OptimizerConfig, model_chunks, no_weight_decay_cond=None,
                           scale_lr_cond=None, lr_mult: float = 1.0):
    if config.optimizer not in ("adam", "adamw"):
        raise ValueError(f"optimizer {config.optimizer!r} is not supported (adam only)")
    if not isinstance(model_chunks, (list, tuple)):
        model_chunks = [model_chunks]
    from ..parallel.fsdp import (FSDPOptimizer, FullyShardedDataParallel, TorchFSDPOptimizer,
                                 TorchFullyShardedDataParallel)

    if all(isinstance(m, FullyShardedDataParallel) for m in model_chunks):
        return FSDPOptimizer(config, model_chunks)
    if all(isinstance(m, TorchFullyShardedDataParallel) for m in model_chunks):
        return TorchFSDPOptimizer(config, model_chunks)
    for m in model_chunks:
        if not isinstance(m, DistributedDataParallel):
            raise TypeError("get_megatron_optimizer expects DistributedDataParallel-wrapped chunks")
    return MegatronOptimizer(config, list(model_chunks), no_weight_decay_cond, scale_lr_cond, lr_mult)


# -- class names of the reference's ``optimizer.py`` ----------------------------------------------------------------------------------
MixedPrecisionOptimizer = MegatronOptimizer          # fp32 master weights + 16-bit model weights are the one optimizer here
Float16OptimizerWithFloat16Params = MegatronOptimizer
DistributedOptimizer = MegatronOptimizer             # ZeRO-1 sharding is a mode of it (``use_distributed_optimizer``)


class ProxyDict:
    """Write-through view on a dict living inside another object (``optimizer.state`` / ``param_groups`` of a chained
    optimizer's members): item access is forwarded to ``getattr(inner, attr)``."""

    def __init__(self, inner_dicts):
        self._inner_dicts = inner_dicts

    def __getitem__(self, key):
        idx, k = key
        return self._inner_dicts[idx][k]

    def __setitem__(self, key, value):
        idx, k = key
        self._inner_dicts[idx][k] = value

    def __len__(self):
        return sum(len(d) for d in self._inner_dicts)

    def __iter__(self):
        for idx, d in enumerate(self._inner_dicts):
            for k in d:
                yield (idx, k)

    def items(self):
        for idx, d in enumerate(self._inner_dicts):
            for k, v in d.items():
                yield (idx, k), v
