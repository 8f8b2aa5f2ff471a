"""FP8 training primitives: device-side tensor-wise quantisation and an fp8 linear layer function whose
three GEMMs (fprop, dgrad, wgrad) run on the block ``tcgen05.mma.kind::f8f6f4`` kernel.

Parity: reference ``core/fp8_utils.py`` (recipes ``tensorwise`` = current scaling and ``delayed`` =
amax history, ``get_fp8_context :384-400``; formats ``e4m3`` / ``hybrid`` = E4M3 forward + E5M2 gradients)
and the TE linear layers it delegates to (``extensions/transformer_engine.py TELinear``).  Everything
(amax, scale, cast, transposed copies, GEMM, dequantisation) stays on the device -- no host sync.

On CPU (tests) the same arithmetic is emulated with ``torch.float8_*`` dtypes.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch

from . import _native, ext, fallback_counts, have_ext

E4M3_MAX, E5M2_MAX = 448.0, 57344.0


@dataclass
class Fp8Recipe:
    fmt: str = "hybrid"            # "e4m3" | "hybrid"
    recipe: str = "tensorwise"     # "tensorwise" (current scaling) | "delayed" | "mxfp8" (1x32 block scaling)
    margin: int = 0
    amax_history_len: int = 1
    amax_compute_algo: str = "most_recent"   # | "max"
    wgrad: bool = True             # False: keep the weight-gradient GEMM in bf16


@dataclass
class Fp8Meta:
    """Per-tensor-role delayed-scaling state (amax history) of one linear layer."""
    history: List[torch.Tensor] = field(default_factory=list)

    def amax_for_scale(self, recipe: Fp8Recipe) -> Optional[torch.Tensor]:
        if not self.history:
            return None
        if recipe.amax_compute_algo == "max":
            return torch.stack(self.history).max(dim=0).values
        return self.history[-1]

    def push(self, amax: torch.Tensor, recipe: Fp8Recipe):
        self.history.append(amax.detach().clone())
        del self.history[:-max(1, recipe.amax_history_len)]


def quantize(x: torch.Tensor, e5m2: bool = False, transpose: bool = False, margin: int = 0,
             amax: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor, torch.Tensor]:
    """``x [rows, cols]`` -> ``(q, q_t, scale_inv, amax)``; ``q`` / ``q_t`` are uint8 storage of the fp8
    values ``x * scale`` with ``scale = fp8_max / amax * 2^-margin`` (``amax`` measured now unless given)."""
    x2 = x.reshape(-1, x.shape[-1])
    if _native(x2) and have_ext() and x2.shape[1] % 8 == 0:
        q, qt, sinv, am = ext().fp8_quantize(x2.contiguous(), e5m2, transpose, margin, amax)
        return q, (qt if transpose else None), sinv, am
    if x2.is_cuda:
        fallback_counts["fp8_quantize"] += 1
    am = x2.detach().abs().max().float().reshape(1) if amax is None else amax
    fmax = E5M2_MAX if e5m2 else E4M3_MAX
    scale = torch.where((am > 0) & torch.isfinite(am), fmax / am * (2.0 ** -margin), torch.ones_like(am))
    dt = torch.float8_e5m2 if e5m2 else torch.float8_e4m3fn
    q = (x2.float() * scale).clamp(-fmax, fmax).to(dt).view(torch.uint8)
    qt = q.t().contiguous() if transpose else None
    return q, qt, (1.0 / scale).float(), am


def dequantize(q: torch.Tensor, scale_inv: torch.Tensor, e5m2: bool = False) -> torch.Tensor:
    dt = torch.float8_e5m2 if e5m2 else torch.float8_e4m3fn
    return q.view(dt).float() * scale_inv


def gemm_fp8(a8, b8, sa_inv, sb_inv, *, a_e5m2=False, b_e5m2=False, bias=None, out=None, out_dtype=torch.bfloat16,
             accumulate=False):
    """``D = sa_inv * sb_inv * A8 @ B8^T (+ bias)``; ``accumulate`` adds into an fp32 ``out``."""
    M, N = a8.shape[0], b8.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32 if accumulate else out_dtype, device=a8.device)
    ok = (a8.is_cuda and have_ext() and a8.shape[1] % 16 == 0 and a8.stride(1) == 1 and b8.stride(1) == 1
          and a8.stride(0) % 16 == 0 and b8.stride(0) % 16 == 0 and out.stride(1) == 1
          and out.stride(0) % (4 if accumulate else 8) == 0)
    if ok:
        ext().gemm_fp8(a8, b8, out, sa_inv, sb_inv, a_e5m2, b_e5m2, bias, accumulate)
        return out
    if a8.is_cuda:
        fallback_counts["gemm_fp8"] += 1
    d = dequantize(a8, sa_inv, a_e5m2) @ dequantize(b8, sb_inv, b_e5m2).t()
    if bias is not None:
        d = d + bias.float()
    if accumulate:
        out.add_(d)
    else:
        out.copy_(d.to(out.dtype))
    return out


def _zero_if_first(weight, main_grad) -> None:
    """The DDP buffer skips zeroing the main_grad of weights whose first wgrad of a step is a plain store
    (``b200_wgrad_store_first``, bf16 path); the fp8 kernels only have the ``+=`` epilogue, so clear it here."""
    if getattr(weight, "b200_wgrad_store_first", False) and not getattr(weight, "grad_added_to_main_grad", False):
        main_grad.zero_()


class _WeightCache:
    """fp8 (and transposed fp8) copies of a weight, rebuilt only when the parameter changes (once per
    optimizer step, not per micro-batch)."""

    def __init__(self):
        self.version, self.data_ptr, self.val = -1, 0, None

    def get(self, w: torch.Tensor, margin: int):
        if self.val is None or self.version != w._version or self.data_ptr != w.data_ptr():
            self.val = quantize(w.detach(), False, True, margin)
            self.version, self.data_ptr = w._version, w.data_ptr()
        return self.val


class Fp8LinearFn(torch.autograd.Function):
    """y = x W^T (+ b) with fp8 fprop / dgrad / wgrad.  ``weight.main_grad`` (fp32) receives the weight
    gradient in place when ``fuse_wgrad`` (same contract as ``ops._LinearFn``)."""

    @staticmethod
    def forward(ctx, x, weight, bias, recipe: Fp8Recipe, cache: _WeightCache, fuse_wgrad: bool):
        x2 = x.reshape(-1, x.shape[-1])
        x8, x8t, xs, _ = quantize(x2, False, True, recipe.margin)
        w8, w8t, ws, _ = cache.get(weight, recipe.margin)
        y = gemm_fp8(x8, w8, xs, ws, bias=bias, out_dtype=x.dtype)
        ctx.save_for_backward(x8t, xs, w8t, ws, x2 if not recipe.wgrad else None)
        ctx.recipe, ctx.fuse_wgrad, ctx.weight = recipe, fuse_wgrad, weight
        ctx.x_shape, ctx.has_bias, ctx.dtype = x.shape, bias is not None, x.dtype
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x8t, xs, w8t, ws, x2 = ctx.saved_tensors
        rec, weight = ctx.recipe, ctx.weight
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        e5 = rec.fmt == "hybrid"
        d8, d8t, ds, _ = quantize(dy2, e5, rec.wgrad, rec.margin)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm_fp8(d8, w8t, ds, ws, a_e5m2=e5, out_dtype=ctx.dtype).view(ctx.x_shape)     # dy @ W
        if ctx.needs_input_grad[1]:
            main = getattr(weight, "main_grad", None)
            if rec.wgrad:
                if ctx.fuse_wgrad and main is not None:
                    _zero_if_first(weight, main)
                    gemm_fp8(d8t, x8t, ds, xs, a_e5m2=e5, out=main, accumulate=True)              # main_grad += dy^T @ x
                    weight.grad_added_to_main_grad = True
                    hook = getattr(weight, "grad_ready_hook", None)
                    if hook is not None and getattr(weight, "b200_wgrad_store_first", False):
                        hook()
                else:
                    dw = gemm_fp8(d8t, x8t, ds, xs, a_e5m2=e5, out_dtype=weight.dtype)
            else:
                from . import wgrad as _wgrad

                dw = _wgrad(dy2, x2, weight, ctx.fuse_wgrad)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            from . import bias_grad

            db = bias_grad(dy2).to(dy.dtype)
        return dx, dw, db, None, None, None


# ---------------------------------------------------------------------------------------------------------------
# MXFP8: OCP micro-scaling -- every 32 consecutive elements along the GEMM's K axis share one power-of-two (E8M0) scale
# ---------------------------------------------------------------------------------------------------------------
MX_BLOCK = 32


def quantize_mx(x: torch.Tensor, e5m2: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """``x [rows, K]`` (K % 32 == 0) -> ``(q uint8 [rows, K], scale_e8m0 uint8 [rows, K/32])``.  The shared exponent
    of a block is ``floor(log2(amax)) - emax_elem`` (OCP MX v1.0 sec. 6.3), stored biased by 127; the elements are
    the saturating fp8 cast of ``x * 2^-shared``."""
    rows, K = x.shape
    if K % MX_BLOCK:
        raise ValueError(f"mxfp8 needs the contraction dim to be a multiple of {MX_BLOCK}, got {K}")
    xb = x.detach().float().view(rows, K // MX_BLOCK, MX_BLOCK)
    amax = xb.abs().amax(dim=-1)
    emax_elem = 15 if e5m2 else 8
    shared = torch.floor(torch.log2(amax.clamp_min(2.0 ** -127))) - emax_elem
    shared = torch.where(amax > 0, shared, torch.full_like(shared, -127.0)).clamp(-127, 127)
    fmax = E5M2_MAX if e5m2 else E4M3_MAX
    dt = torch.float8_e5m2 if e5m2 else torch.float8_e4m3fn
    q = (xb * torch.exp2(-shared).unsqueeze(-1)).clamp(-fmax, fmax).to(dt).view(torch.uint8).view(rows, K)
    return q, (shared + 127).to(torch.uint8)


def dequantize_mx(q: torch.Tensor, scale_e8m0: torch.Tensor, e5m2: bool = False, dtype=torch.float32) -> torch.Tensor:
    dt = torch.float8_e5m2 if e5m2 else torch.float8_e4m3fn
    rows, K = q.shape
    s = torch.exp2(scale_e8m0.float() - 127.0).unsqueeze(-1)
    return (q.view(dt).float().view(rows, K // MX_BLOCK, MX_BLOCK) * s).view(rows, K).to(dtype)


def gemm_mx(a8, sa, b8, sb, *, a_e5m2=False, b_e5m2=False, bias=None, out_dtype=torch.bfloat16):
    """``D = deq(A) @ deq(B)^T``; both operands are block-scaled along K.  The block scales are exact powers of two,
    so dequantising to bf16 is lossless (fp8 mantissas fit) and the product runs on the bf16 tcgen05 GEMM; a native
    ``tcgen05.mma.kind::mxf8f6f4.block_scale`` kernel only changes the speed, not the numbers."""
    if a8.is_cuda:
        from . import gemm as _gemm

        fallback_counts["mxfp8_via_bf16_gemm"] += 1
        d = _gemm(dequantize_mx(a8, sa, a_e5m2, torch.bfloat16), dequantize_mx(b8, sb, b_e5m2, torch.bfloat16))
        d = d.float()
    else:
        d = dequantize_mx(a8, sa, a_e5m2) @ dequantize_mx(b8, sb, b_e5m2).t()
    if bias is not None:
        d = d + bias.float()
    return d.to(out_dtype)


# ---- native path: tcgen05.mma.kind::mxf8f6f4.block_scale (csrc/gemm_mxfp8_sm100.cu) ---------------------------------------
MX_BN = 192        # N-tile of the native kernel = height of a B-operand scale tile


@dataclass
class MxTensor:
    """fp8 payload ``q [rows, K]`` + E8M0 scales in the kernel's chunk layout (``role`` 0 = A operand, 1 = B operand)."""
    q: torch.Tensor
    sf: torch.Tensor
    role: int
    e5m2: bool = False


def mx_native_ok(t: torch.Tensor) -> bool:
    return _native(t) and have_ext()


def quantize_mx_native(x2: torch.Tensor, e5m2: bool, role_row: int = -1, role_col: int = -1):
    """One kernel: row-wise copy (blocks along the last dim) and / or column-wise copy (x^T, blocks along the rows)."""
    q_row, sf_row, q_col, sf_col = ext().mx_quantize(x2.contiguous(), e5m2, role_row, role_col)
    return (MxTensor(q_row, sf_row, role_row, e5m2) if role_row >= 0 else None,
            MxTensor(q_col, sf_col, role_col, e5m2) if role_col >= 0 else None)


def sf_matrix(t: MxTensor) -> torch.Tensor:
    """Scales of ``t`` as a plain ``[rows, K/32]`` uint8 matrix (de-layout of the chunked storage; tests / debugging)."""
    rows, K = t.q.shape
    r = torch.arange(rows, device=t.sf.device).unsqueeze(1)
    kb32 = torch.arange(K // 32, device=t.sf.device).unsqueeze(0)
    kb, ks, kblocks = kb32 // 4, kb32 % 4, K // 128
    if t.role == 0:
        chunk, rr = (r // 128) * kblocks + kb, r % 128
    else:
        tile, r2 = r // MX_BN, r % MX_BN
        chunk, rr = (tile * kblocks + kb) * 2 + r2 // 128, r2 % 128
    off = chunk * 512 + (rr % 32) * 16 + (rr // 32) * 4 + ks
    return t.sf[off.reshape(-1).long()].view(rows, K // 32)


def dequantize_mx_native(t: MxTensor, dtype=torch.float32) -> torch.Tensor:
    return dequantize_mx(t.q, sf_matrix(t), t.e5m2, dtype)


def gemm_mx_native(a: MxTensor, b: MxTensor, *, bias=None, out=None, out_dtype=torch.bfloat16, accumulate=False):
    """``D = deq(a) @ deq(b)^T`` on the block-scaled tensor-core path (scales applied inside the MMA)."""
    assert a.role == 0 and b.role == 1
    M, N = a.q.shape[0], b.q.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32 if accumulate else out_dtype, device=a.q.device)
    ext().gemm_mxfp8(a.q, a.sf, b.q, b.sf, out, a.e5m2, b.e5m2, bias, accumulate)
    return out


class _MxWeightCache(_WeightCache):
    """Row-wise (blocks along ``in``: fprop) and column-wise (blocks along ``out``: dgrad) mxfp8 copies of a weight."""

    def get(self, w: torch.Tensor, margin: int = 0):
        if self.val is None or self.version != w._version or self.data_ptr != w.data_ptr():
            wd = w.detach()
            if mx_native_ok(wd) and wd.shape[1] % 128 == 0 and wd.shape[0] % 128 == 0:
                self.val = ("native",) + quantize_mx_native(wd, False, 1, 1)     # (W, blocks along in) , (W^T, blocks along out)
            else:
                ok_t = wd.shape[0] % MX_BLOCK == 0
                self.val = (*quantize_mx(wd), *(quantize_mx(wd.t().contiguous()) if ok_t else (None, None)))
            self.version, self.data_ptr = w._version, w.data_ptr()
        return self.val


class MxFp8LinearNativeFn(torch.autograd.Function):
    """mxfp8 linear on the native block-scaled GEMM: every operand is quantised ONCE per use-direction by a dual
    (row-wise + column-wise) kernel, the three GEMMs read fp8 + E8M0 scales directly -- no bf16 dequantisation."""

    @staticmethod
    def forward(ctx, x, weight, bias, recipe: Fp8Recipe, cache: _MxWeightCache, fuse_wgrad: bool):
        x2 = x.reshape(-1, x.shape[-1])
        rows = x2.shape[0]
        want_wgrad = weight.requires_grad and recipe.wgrad and rows % 128 == 0
        xr, xc = quantize_mx_native(x2, False, 0, 1 if want_wgrad else -1)       # X (A of fprop), X^T (B of wgrad)
        _, wr, wc = cache.get(weight)
        y = gemm_mx_native(xr, wr, bias=bias, out_dtype=x.dtype)
        ctx.xc, ctx.wc = xc, wc
        ctx.save_for_backward(x2 if xc is None else None)
        ctx.recipe, ctx.fuse_wgrad, ctx.weight = recipe, fuse_wgrad, weight
        ctx.x_shape, ctx.has_bias, ctx.dtype = x.shape, bias is not None, x.dtype
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        (x2,) = ctx.saved_tensors
        rec, weight = ctx.recipe, ctx.weight
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        e5 = rec.fmt == "hybrid"
        dr, dc = quantize_mx_native(dy2, e5, 0, 0 if ctx.xc is not None else -1)  # dY (A of dgrad), dY^T (A of wgrad)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm_mx_native(dr, ctx.wc, out_dtype=ctx.dtype).view(ctx.x_shape)          # dY @ W
        if ctx.needs_input_grad[1]:
            main = getattr(weight, "main_grad", None)
            if ctx.xc is not None:
                if ctx.fuse_wgrad and main is not None and main.dtype == torch.float32:
                    _zero_if_first(weight, main)
                    gemm_mx_native(dc, ctx.xc, out=main, accumulate=True)                    # main_grad += dY^T @ X
                    weight.grad_added_to_main_grad = True
                    hook = getattr(weight, "grad_ready_hook", None)
                    if hook is not None and getattr(weight, "b200_wgrad_store_first", False):
                        hook()
                else:
                    dw = gemm_mx_native(dc, ctx.xc, out_dtype=weight.dtype)
            else:
                from . import wgrad as _wgrad

                dw = _wgrad(dy2, x2, weight, ctx.fuse_wgrad)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            from . import bias_grad

            db = bias_grad(dy2).to(dy.dtype)
        ctx.xc = ctx.wc = None
        return dx, dw, db, None, None, None


class MxFp8LinearFn(torch.autograd.Function):
    """Linear layer under the ``mxfp8`` recipe (reference ``fp8_utils.get_fp8_context``: ``MXFP8BlockScaling``).
    Each GEMM quantises its two operands along ITS contraction axis: fprop along ``in``, dgrad along ``out``, wgrad
    along tokens -- so activations and gradients are quantised twice (row- and column-wise), like TE does."""

    @staticmethod
    def forward(ctx, x, weight, bias, recipe: Fp8Recipe, cache: _MxWeightCache, fuse_wgrad: bool):
        x2 = x.reshape(-1, x.shape[-1])
        xq, xs = quantize_mx(x2)
        val = cache.get(weight)
        if val[0] == "native":            # cached by the native path for another input shape: rebuild in matrix form
            wd = weight.detach()
            val = (*quantize_mx(wd), *(quantize_mx(wd.t().contiguous()) if wd.shape[0] % MX_BLOCK == 0 else (None, None)))
        wq, ws, wtq, wts = val
        y = gemm_mx(xq, xs, wq, ws, bias=bias, out_dtype=x.dtype)
        ctx.save_for_backward(x2, wtq, wts)
        ctx.recipe, ctx.fuse_wgrad, ctx.weight = recipe, fuse_wgrad, weight
        ctx.x_shape, ctx.has_bias, ctx.dtype = x.shape, bias is not None, x.dtype
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, wtq, wts = ctx.saved_tensors
        rec, weight = ctx.recipe, ctx.weight
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        e5 = rec.fmt == "hybrid"
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if wtq is not None:
                dq, ds = quantize_mx(dy2, e5)                                        # blocks along ``out``
                dx = gemm_mx(dq, ds, wtq, wts, a_e5m2=e5, out_dtype=ctx.dtype)       # dy @ W
            else:
                dx = (dy2.float() @ weight.detach().float()).to(ctx.dtype)
            dx = dx.view(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            rows = dy2.shape[0]
            if rec.wgrad and rows % MX_BLOCK == 0:
                dtq, dts = quantize_mx(dy2.t().contiguous(), e5)                     # blocks along tokens
                xtq, xts = quantize_mx(x2.t().contiguous())
                dw = gemm_mx(dtq, dts, xtq, xts, a_e5m2=e5, out_dtype=torch.float32)  # dy^T @ x
            else:
                dw = dy2.float().t() @ x2.float()
            main = getattr(weight, "main_grad", None)
            if ctx.fuse_wgrad and main is not None:
                _zero_if_first(weight, main)
                main.add_(dw)
                weight.grad_added_to_main_grad = True
                dw = None
            else:
                dw = dw.to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            from . import bias_grad

            db = bias_grad(dy2).to(dy.dtype)
        return dx, dw, db, None, None, None


def fp8_linear(x, weight, bias=None, recipe: Optional[Fp8Recipe] = None, cache: Optional[_WeightCache] = None,
               fuse_wgrad: bool = False):
    recipe = recipe or Fp8Recipe()
    if recipe.recipe == "mxfp8" and x.shape[-1] % MX_BLOCK == 0:
        if not isinstance(cache, _MxWeightCache):
            cache = _MxWeightCache()
        if (mx_native_ok(x) and x.shape[-1] % 128 == 0 and weight.shape[0] % 128 == 0 and weight.dtype == x.dtype
                and x.reshape(-1, x.shape[-1]).shape[0] > 0):
            return MxFp8LinearNativeFn.apply(x, weight, bias, recipe, cache, fuse_wgrad)
        return MxFp8LinearFn.apply(x, weight, bias, recipe, cache, fuse_wgrad)
    return Fp8LinearFn.apply(x, weight, bias, recipe or Fp8Recipe(), cache or _WeightCache(), fuse_wgrad)


# -- names of the reference's ``core/fp8_utils.py`` (Transformer-Engine Float8Tensor parameters do not exist here) ----------------
def is_float8tensor(tensor) -> bool:
    """Parameters are never stored as fp8 tensors: weights are quantised per step into a cache (``_WeightCache``)."""
    return False


def modify_underlying_storage(tensor, new_raw_data):
    tensor.data = new_raw_data


def quantize_param_shard(model_params, main_params, start_offsets, data_parallel_group, fsdp_shard_model_params=None):
    """fp8 parameter shards (``--fp8-param-gather``) are not used: the bf16 parameters are updated by the optimizer directly."""
    return None


def correct_amax_history_if_needed(model):
    """Delayed-scaling amax histories live in ``Fp8Meta`` objects next to each linear and need no fix-up after a load."""
    return None
