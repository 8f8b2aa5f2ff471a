"""Post-training quantisation (PTQ) and quantisation-aware fine-tuning hooks.

Parity: the reference delegates to NVIDIA ModelOpt (``megatron/post_training/{model_provider,checkpointing}.py``,
``core/post_training/modelopt/gpt/{model_specs,state_dict_hooks}.py``, ``examples/post_training/modelopt/quantize.py``:
``mtq.quantize(model, QUANT_CFG_CHOICES[name], forward_loop)`` then a checkpoint that carries the quantiser state).
ModelOpt is not part of this stack, so the same workflow is implemented here:

    cfg = QUANT_CFG_CHOICES["fp8"]
    quantize(model, cfg, forward_loop)          # insert quantisers -> calibrate (max) -> fold / enable fake quant
    sd = quantizer_state_dict(model)            # amax buffers, saved next to the weights
    real = export_quantized_weights(model)      # int8 / fp8 / int4 payloads + scales for an inference runtime

Formats: ``fp8`` (per-tensor E4M3 activations and weights), ``int8_sq`` (SmoothQuant: per-channel int8 weights,
per-tensor int8 activations, the smoothing vector kept as a pre-quant scale like TRT-LLM's ``prequant_scaling_factor``), ``int4_awq`` (weight-only, group
wise int4 with an activation-aware per-channel scale search), ``nvfp4`` (E2M1 weights in blocks of 16 with an fp8
block scale) and ``w4a8_awq``.  Calibration is tensor-parallel aware: per-tensor amax is MAX-reduced over the TP
group (a sharded linear must use one activation scale), per-channel statistics stay with their shard.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Iterable, Optional, Tuple

import torch
import torch.nn as nn

E4M3_MAX = 448.0
_FP4_GRID = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0])


@dataclass(frozen=True)
class QuantFormat:
    kind: str                     # "int" | "fp8" | "fp4"
    bits: int = 8
    axis: Optional[int] = None    # per-channel axis (weights: 0 = output channel); None = per-tensor
    block: Optional[int] = None   # group size along the input dim (int4 AWQ: 128, nvfp4: 16)


@dataclass(frozen=True)
class QuantConfig:
    name: str
    weight: Optional[QuantFormat]
    activation: Optional[QuantFormat]
    algorithm: str = "max"        # "max" | "smoothquant" | "awq_lite"
    alpha: float = 0.5            # SmoothQuant migration strength


QUANT_CFG_CHOICES: Dict[str, QuantConfig] = {
    "fp8": QuantConfig("fp8", QuantFormat("fp8"), QuantFormat("fp8")),
    "int8_sq": QuantConfig("int8_sq", QuantFormat("int", 8, axis=0), QuantFormat("int", 8), "smoothquant"),
    "int4_awq": QuantConfig("int4_awq", QuantFormat("int", 4, block=128), None, "awq_lite"),
    "w4a8_awq": QuantConfig("w4a8_awq", QuantFormat("int", 4, block=128), QuantFormat("fp8"), "awq_lite"),
    "nvfp4": QuantConfig("nvfp4", QuantFormat("fp4", 4, block=16), QuantFormat("fp4", 4, block=16)),
}


# ---------------------------------------------------------------------------------------------------------------
# fake quantisation (quantise -> dequantise), straight-through in backward
# ---------------------------------------------------------------------------------------------------------------
def _blocked(x: torch.Tensor, block: int) -> torch.Tensor:
    if x.shape[-1] % block:
        raise ValueError(f"last dim {x.shape[-1]} is not a multiple of the quantisation block {block}")
    return x.reshape(*x.shape[:-1], x.shape[-1] // block, block)


def _amax(x: torch.Tensor, fmt: QuantFormat) -> torch.Tensor:
    x = x.detach().float().abs()
    if fmt.block:
        return _blocked(x, fmt.block).amax(-1)
    if fmt.axis is not None:
        dims = [d for d in range(x.dim()) if d != fmt.axis % x.dim()]
        return x.amax(dims) if dims else x
    return x.amax().reshape(())


def _expand(amax: torch.Tensor, x: torch.Tensor, fmt: QuantFormat) -> torch.Tensor:
    if fmt.block:
        return amax.unsqueeze(-1)
    if fmt.axis is not None:
        shape = [1] * x.dim()
        shape[fmt.axis % x.dim()] = -1
        return amax.view(shape)
    return amax


def fake_quant(x: torch.Tensor, amax: torch.Tensor, fmt: QuantFormat) -> torch.Tensor:
    xf = _blocked(x.float(), fmt.block) if fmt.block else x.float()
    a = _expand(amax.float().to(x.device), x, fmt).clamp_min(1e-12)
    if fmt.kind == "int":
        qmax = 2 ** (fmt.bits - 1) - 1
        y = torch.round(xf / a * qmax).clamp(-qmax - 1, qmax) * (a / qmax)
    elif fmt.kind == "fp8":
        s = E4M3_MAX / a
        y = (xf * s).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).float() / s
    elif fmt.kind == "fp4":
        # block scale itself is stored in fp8 (E4M3), as in NVFP4
        s = (a / 6.0).to(torch.float8_e4m3fn).float().clamp_min(2.0 ** -9)
        grid = _FP4_GRID.to(x.device)
        mag = (xf.abs() / s).clamp(max=6.0)
        idx = (mag.unsqueeze(-1) - grid).abs().argmin(-1)
        y = grid[idx] * xf.sign() * s
    else:
        raise ValueError(fmt.kind)
    return y.reshape(x.shape).to(x.dtype)


class _STE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, amax, fmt):
        return fake_quant(x, amax, fmt)

    @staticmethod
    def backward(ctx, g):
        return g, None, None


class TensorQuantizer(nn.Module):
    """Calibrate-then-quantise node; ``_amax`` is a buffer so it travels with ``state_dict``."""

    def __init__(self, fmt: QuantFormat, tp_group=None, dynamic: bool = False):
        super().__init__()
        self.fmt, self.tp_group, self.dynamic = fmt, tp_group, dynamic      # dynamic: amax measured per call
        self.calibrating, self.enabled = False, False
        self.register_buffer("_amax", None)

    @property
    def amax(self):
        return self._amax

    def observe(self, x: torch.Tensor):
        a = _amax(x, self.fmt)
        self._amax = a if self._amax is None else torch.maximum(self._amax, a)

    def finish_calibration(self):
        self.calibrating = False
        if self.dynamic:
            self._amax, self.enabled = None, True
            return
        if self._amax is not None and self.fmt.axis is None and self.fmt.block is None and self.tp_group is not None \
                and torch.distributed.is_initialized() and torch.distributed.get_world_size(self.tp_group) > 1:
            torch.distributed.all_reduce(self._amax, op=torch.distributed.ReduceOp.MAX, group=self.tp_group)
        self.enabled = self._amax is not None

    def forward(self, x):
        if self.calibrating:
            if not self.dynamic:
                self.observe(x)
            return x
        if not self.enabled:
            return x
        return _STE.apply(x, _amax(x, self.fmt) if self.dynamic else self._amax, self.fmt)

    def extra_repr(self):
        return f"{self.fmt}, enabled={self.enabled}"


# ---------------------------------------------------------------------------------------------------------------
# model surgery
# ---------------------------------------------------------------------------------------------------------------
def _quantizable(model: nn.Module, skip: Iterable[str]) -> Iterable[Tuple[str, nn.Module]]:
    from ..parallel.tp_layers import ColumnParallelLinear, RowParallelLinear

    for name, m in model.named_modules():
        if isinstance(m, (ColumnParallelLinear, RowParallelLinear, nn.Linear)) and getattr(m, "weight", None) is not None \
                and not any(s in name for s in skip):
            yield name, m


def _tp_group():
    from ..parallel import state as ps

    return ps.get_tensor_model_parallel_group() if ps.model_parallel_is_initialized() else None


def _pre_hook(mod, args):
    q = mod.input_quantizer
    if mod._ptq_smooth is not None:
        args = (args[0] / mod._ptq_smooth.to(args[0].dtype),) + tuple(args[1:])
    if mod._ptq_stats is not None:                      # per-input-channel activation magnitude (SmoothQuant / AWQ)
        a = args[0].detach().float().abs().reshape(-1, args[0].shape[-1]).amax(0)
        mod._ptq_stats = a if mod._ptq_stats.numel() == 0 else torch.maximum(mod._ptq_stats, a)
    return (q(args[0]),) + tuple(args[1:]) if q is not None else args


def insert_quantizers(model: nn.Module, cfg: QuantConfig, skip: Iterable[str] = ("output_layer", "router")):
    """Attach ``input_quantizer`` / ``weight_quantizer`` to every linear layer (embedding, LM head and MoE routers
    stay in 16 bit, like ModelOpt's default configs)."""
    n = 0
    for _, m in _quantizable(model, skip):
        if hasattr(m, "weight_quantizer"):
            continue
        row = type(m).__name__ == "RowParallelLinear"
        m.input_quantizer = (TensorQuantizer(cfg.activation, _tp_group() if row else None,
                                             dynamic=cfg.activation.block is not None) if cfg.activation else None)
        m.weight_quantizer = TensorQuantizer(cfg.weight) if cfg.weight else None
        m._ptq_smooth, m._ptq_stats = None, None
        m._ptq_hook = m.register_forward_pre_hook(_pre_hook)
        n += 1
    return n


def _search_awq_scale(w: torch.Tensor, act: torch.Tensor, fmt: QuantFormat, grid: int = 10) -> torch.Tensor:
    """AWQ-lite: per-input-channel scale ``s = act^a`` (a on a grid) minimising the activation-weighted weight error."""
    act = act.to(w.device).clamp_min(1e-5)
    best, best_err = torch.ones_like(act), None
    wf = w.float()
    for i in range(grid + 1):
        s = act.pow(i / grid)
        s = s / (s.max() * s.min()).sqrt()
        wq = fake_quant(wf * s, _amax(wf * s, fmt), fmt) / s
        err = (((wq - wf) * act) ** 2).sum()
        if best_err is None or err < best_err:
            best, best_err = s, err
    return best


@torch.no_grad()
def quantize(model: nn.Module, cfg, forward_loop: Optional[Callable[[nn.Module], None]] = None,
             fold_weights: bool = True, skip: Iterable[str] = ("output_layer", "router")) -> nn.Module:
    """ModelOpt-style entry point: insert quantisers, run ``forward_loop(model)`` for calibration, then switch the
    quantisers on.  With ``fold_weights`` the weights are replaced by their quantise->dequantise image once (PTQ);
    otherwise the weight quantiser is applied on every forward with a straight-through gradient (QAT)."""
    if isinstance(cfg, str):
        cfg = QUANT_CFG_CHOICES[cfg]
    insert_quantizers(model, cfg, skip)
    layers = [m for _, m in _quantizable(model, skip) if hasattr(m, "weight_quantizer")]
    need_stats = cfg.algorithm in ("smoothquant", "awq_lite")
    for m in layers:
        if m.input_quantizer is not None:
            m.input_quantizer.calibrating = True
        m._ptq_stats = torch.empty(0) if need_stats else None
    if forward_loop is not None:
        was = model.training
        model.eval()
        forward_loop(model)
        model.train(was)
    elif cfg.activation is not None or need_stats:
        raise ValueError(f"{cfg.name} needs a calibration forward_loop")
    for m in layers:
        stats = m._ptq_stats if need_stats and m._ptq_stats is not None and m._ptq_stats.numel() else None
        m._ptq_stats = None
        w = m.weight
        if stats is not None and cfg.algorithm == "smoothquant":
            s = (stats.to(w.device).clamp_min(1e-5).pow(cfg.alpha)
                 / w.detach().float().abs().amax(0).clamp_min(1e-5).pow(1 - cfg.alpha)).clamp(1e-4, 1e4)
            w.mul_(s.to(w.dtype))                      # W' = W diag(s),  x' = x / s
            m._ptq_smooth = s
            if m.input_quantizer is not None and m.input_quantizer._amax is not None:
                m.input_quantizer._amax = (stats.to(w.device) / s).amax().reshape(())
        elif stats is not None and cfg.algorithm == "awq_lite":
            s = _search_awq_scale(w.detach(), stats, cfg.weight)
            w.mul_(s.to(w.dtype))
            m._ptq_smooth = s
            if m.input_quantizer is not None and m.input_quantizer._amax is not None:
                m.input_quantizer._amax = (stats.to(w.device) / s).amax().reshape(())
        if m.input_quantizer is not None:
            m.input_quantizer.finish_calibration()
        if m.weight_quantizer is not None:
            m.weight_quantizer.observe(w)
            m.weight_quantizer.finish_calibration()
            if fold_weights:
                w.copy_(fake_quant(w, m.weight_quantizer._amax, cfg.weight))
                m.weight_quantizer.enabled = False      # already on the grid
            else:
                _install_weight_fake_quant(m)
    model._ptq_config = cfg
    return model


def _install_weight_fake_quant(m: nn.Module):
    """QAT: quantise the weight on the fly (straight-through).  ``ColumnParallelLinear.forward`` takes ``weight=``;
    other layers get a parametrisation-free swap around their forward."""
    orig = m.forward

    def fwd(*a, **k):
        wq = m.weight_quantizer(m.weight)
        if "weight" in orig.__code__.co_varnames[: orig.__code__.co_argcount + 1]:
            return orig(*a, weight=wq, **k)
        saved = m.weight.data
        m.weight.data = wq.detach()
        try:
            return orig(*a, **k)
        finally:
            m.weight.data = saved

    m.forward = fwd


# ---------------------------------------------------------------------------------------------------------------
# state / export
# ---------------------------------------------------------------------------------------------------------------
def quantizer_state_dict(model: nn.Module) -> Dict[str, torch.Tensor]:
    out = {}
    for name, m in model.named_modules():
        if isinstance(m, TensorQuantizer) and m._amax is not None:
            out[f"{name}._amax"] = m._amax.detach().clone()
    for name, m in model.named_modules():
        if getattr(m, "_ptq_smooth", None) is not None:
            out[f"{name}._pre_quant_scale"] = m._ptq_smooth.detach().clone()
    return out


def restore_quantizers(model: nn.Module, cfg, state: Dict[str, torch.Tensor]):
    """Rebuild the quantised model from a checkpoint: insert quantisers, load amax / smoothing vectors, enable."""
    if isinstance(cfg, str):
        cfg = QUANT_CFG_CHOICES[cfg]
    insert_quantizers(model, cfg)
    mods = dict(model.named_modules())
    for k, v in state.items():
        name, leaf = k.rsplit(".", 1)
        if leaf == "_amax":
            mods[name]._amax = v.clone()
            mods[name].enabled = "weight_quantizer" not in name      # folded weights are already on the grid
        elif leaf == "_pre_quant_scale":
            mods[name]._ptq_smooth = v.clone()
    for m in mods.values():
        if isinstance(m, TensorQuantizer) and m.dynamic:
            m.enabled = True
    model._ptq_config = cfg
    return model


@torch.no_grad()
def export_quantized_weights(model: nn.Module) -> Dict[str, torch.Tensor]:
    """Real low-precision payloads for an inference runtime: ``<layer>.weight`` (int8 / fp8-as-uint8 / packed int4),
    ``<layer>.weights_scaling_factor`` and, when activations are quantised, ``<layer>.activation_scaling_factor``
    (+ ``prequant_scaling_factor`` for SmoothQuant / AWQ) -- the TRT-LLM checkpoint field names."""
    cfg: QuantConfig = model._ptq_config
    out = {}
    for name, m in model.named_modules():
        wq = getattr(m, "weight_quantizer", None)
        if wq is None or wq._amax is None:
            continue
        fmt, w, a = wq.fmt, m.weight.detach().float(), wq._amax.float()
        if fmt.kind == "fp8":
            s = a / E4M3_MAX
            out[f"{name}.weight"] = (w / s).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).view(torch.uint8)
        elif fmt.kind == "int":
            qmax = 2 ** (fmt.bits - 1) - 1
            s = _expand(a, w, fmt) / qmax
            wb = _blocked(w, fmt.block) if fmt.block else w
            q = torch.round(wb / s.clamp_min(1e-12)).clamp(-qmax - 1, qmax).to(torch.int8).reshape(w.shape)
            if fmt.bits == 4:                           # two nibbles per byte, low nibble first
                q = ((q[..., 0::2] & 0xF) | ((q[..., 1::2] & 0xF) << 4)).to(torch.uint8)
            out[f"{name}.weight"] = q
            s = s.squeeze(-1) if fmt.block else s.reshape(-1)
        else:                                           # fp4: store the dequantised grid value index + fp8 block scale
            s = (a / 6.0).to(torch.float8_e4m3fn)
            grid = _FP4_GRID
            mag = (_blocked(w, fmt.block).abs() / s.float().clamp_min(2.0 ** -9).unsqueeze(-1)).clamp(max=6.0)
            idx = (mag.unsqueeze(-1) - grid).abs().argmin(-1).to(torch.uint8).reshape(w.shape)
            idx = idx | ((w < 0).to(torch.uint8) << 3)
            out[f"{name}.weight"] = (idx[..., 0::2] | (idx[..., 1::2] << 4)).to(torch.uint8)
            s = s.view(torch.uint8)
        out[f"{name}.weights_scaling_factor"] = s
        iq = getattr(m, "input_quantizer", None)
        if iq is not None and iq._amax is not None:
            den = E4M3_MAX if iq.fmt.kind == "fp8" else (6.0 if iq.fmt.kind == "fp4" else 2 ** (iq.fmt.bits - 1) - 1)
            out[f"{name}.activation_scaling_factor"] = iq._amax.float() / den
        if getattr(m, "_ptq_smooth", None) is not None:
            out[f"{name}.prequant_scaling_factor"] = 1.0 / m._ptq_smooth.float()
    return out
