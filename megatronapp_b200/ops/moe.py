"""MoE operator layer: ragged (256-row padded) expert layout, grouped tcgen05 GEMMs, fused combine.

CUDA path = ``csrc/moe.cu`` + the grouped modes of ``csrc/gemm2_sm100.cu`` -- no host synchronisation anywhere:
the layout (per-expert offsets, row <-> (token, expert) maps) is built and consumed on the device.  The PyTorch
path below implements the *same layout* with index ops and is what the CPU tests compare against
``SequentialMLP`` (reference ``transformer/moe/experts.py:90-265``, ``moe_utils.py`` permute/unpermute).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import _native, ext, gemm

PAD = 256   # CTA-pair tile height of the grouped GEMM


@dataclass
class RaggedLayout:
    """Device-resident description of ``[capacity, hidden]`` expert input/output buffers."""
    n_tokens: int
    E_total: int
    e_begin: int
    E_local: int
    cap: int
    counts: torch.Tensor      # int32 [E_local]
    offsets: torch.Tensor     # int32 [E_local + 1], multiples of PAD
    row_slot: torch.Tensor    # int32 [cap]: token * E_total + expert, -1 for padding rows
    row_prob: torch.Tensor    # fp32 [cap]
    dest_row: torch.Tensor    # int32 [n_tokens, E_local]: row of (token, local expert) or -1

    @property
    def tokens_per_expert(self) -> torch.Tensor:
        return self.counts


def capacity_for(n_tokens: int, max_copies_per_token: int, E_local: int, pad: int = PAD) -> int:
    k = min(max_copies_per_token, E_local)
    return (n_tokens * k + E_local * (pad - 1) + pad - 1) // pad * pad


def build_layout(routing_map: torch.Tensor, probs: Optional[torch.Tensor], e_begin: int, E_local: int, topk: int,
                 pad: int = PAD) -> RaggedLayout:
    n, E = routing_map.shape
    cap = capacity_for(n, topk, E_local, pad)
    if routing_map.is_cuda:
        p = probs.detach().float().contiguous() if probs is not None else None
        counts, offsets, row_slot, row_prob, dest_row = ext().moe_build_layout(routing_map.contiguous(), p, e_begin, E_local,
                                                                               pad, cap)
        return RaggedLayout(n, E, e_begin, E_local, cap, counts, offsets, row_slot, row_prob, dest_row)
    # reference path (same layout)
    rm = routing_map[:, e_begin:e_begin + E_local].bool()
    counts = rm.sum(0).to(torch.int32)
    padded = (counts + pad - 1) // pad * pad
    offsets = torch.zeros(E_local + 1, dtype=torch.int32)
    offsets[1:] = torch.cumsum(padded, 0)
    pos = torch.cumsum(rm.int(), 0) - 1
    dest_row = torch.where(rm, pos + offsets[:-1].unsqueeze(0), torch.full_like(pos, -1)).to(torch.int32)
    row_slot = torch.full((cap,), -1, dtype=torch.int32)
    row_prob = torch.zeros(cap, dtype=torch.float32)
    t_idx, e_idx = rm.nonzero(as_tuple=True)
    rows = dest_row[t_idx, e_idx].long()
    row_slot[rows] = (t_idx * E + e_begin + e_idx).to(torch.int32)
    if probs is not None:
        row_prob[rows] = probs.detach().float()[t_idx, e_begin + e_idx]
    else:
        row_prob[rows] = 1.0
    return RaggedLayout(n, E, e_begin, E_local, cap, counts, offsets, row_slot, row_prob, dest_row)


# ---------------------------------------------------------------------------------------------- row movers
def _gather_rows(x, lay: RaggedLayout, slot_div: int, scale=None, other=None, dot_out=None):
    if _native(x) and x.shape[1] % 8 == 0:
        return ext().moe_gather_rows(x.contiguous(), lay.row_slot, slot_div, scale, lay.offsets, lay.E_local, other, dot_out)
    valid = lay.row_slot >= 0
    src = (lay.row_slot.clamp(min=0) // slot_div).long()
    out = x.float()[src]
    if dot_out is not None:
        d = (out * other.float()).sum(-1)
        d = torch.where(scale != 0, d / scale, torch.zeros_like(d))
        dot_out.view(-1)[lay.row_slot[valid].long()] = d[valid]
    if scale is not None:
        out = out * scale.unsqueeze(-1)
    out = out * valid.unsqueeze(-1)
    return out.to(x.dtype)


def _gather_sum(y, lay: RaggedLayout):
    if _native(y) and y.shape[1] % 8 == 0:
        return ext().moe_gather_sum(y.contiguous(), lay.dest_row)
    dr = lay.dest_row.long()
    g = y.float()[dr.clamp(min=0)] * (dr >= 0).unsqueeze(-1)
    return g.sum(1).to(y.dtype)


class _Permute(torch.autograd.Function):
    """tokens [n, h] -> ragged expert rows [cap, h] (zero padding rows); backward sums the copies of each token."""

    @staticmethod
    def forward(ctx, x, lay: RaggedLayout):
        ctx.lay = lay
        return _gather_rows(x, lay, lay.E_total)

    @staticmethod
    def backward(ctx, dxp):
        return _gather_sum(dxp.contiguous(), ctx.lay), None


def permute(x: torch.Tensor, lay: RaggedLayout) -> torch.Tensor:
    return _Permute.apply(x, lay)


# ---------------------------------------------------------------------------------------------- grouped GEMM
def _grouped_ref(a, w, lay: RaggedLayout, stride: int, b_mn: bool, row_scale=None):
    """D[rows of g] = A[rows of g] @ W_g^T (or @ W_g when the weight is consumed MN-major)."""
    offs = lay.offsets.tolist()
    n_out = w.shape[1] if b_mn else stride
    out = torch.zeros(a.shape[0], n_out, dtype=torch.float32, device=a.device)
    for g in range(lay.E_local):
        r0, r1 = offs[g], offs[g + 1]
        if r1 > r0:
            wg = w[g * stride:(g + 1) * stride].float()
            out[r0:r1] = a[r0:r1].float() @ (wg if b_mn else wg.t())
    if row_scale is not None:
        out = out * row_scale.unsqueeze(-1)
    return out.to(a.dtype)


def grouped_gemm(a, w, lay: RaggedLayout, stride: int, b_mn: bool = False, row_scale=None):
    if _native(a) and a.dtype == w.dtype and a.shape[1] % 8 == 0 and w.shape[1] % 8 == 0:
        n_out = w.shape[1] if b_mn else stride
        out = torch.empty(a.shape[0], n_out, dtype=a.dtype, device=a.device)
        ext().gemm_grouped(a, w, out, lay.offsets, 1, b_mn, None, 0, None, row_scale, lay.E_local, stride)
        return out
    return _grouped_ref(a, w, lay, stride, b_mn, row_scale)


def grouped_wgrad(dy, x, weight, lay: RaggedLayout, stride: int, fuse: bool):
    """dW_g = dY_g^T X_g for every group; accumulates into ``weight.main_grad`` when fused (returns None then)."""
    main_grad = getattr(weight, "main_grad", None)
    native = _native(dy) and dy.dtype == x.dtype and stride % 256 == 0 and x.shape[1] % 8 == 0

    def run(dst32):
        if native:
            ext().gemm_grouped(dy, x, dst32, lay.offsets, 2, True, None, 0, None, None, lay.E_local, stride)
        else:
            offs = lay.offsets.tolist()
            for g in range(lay.E_local):
                r0, r1 = offs[g], offs[g + 1]
                if r1 > r0:
                    dst32[g * stride:(g + 1) * stride] += dy[r0:r1].float().t() @ x[r0:r1].float()

    if fuse and main_grad is not None and main_grad.dtype == torch.float32:
        run(main_grad)
        weight.grad_added_to_main_grad = True
        hook = getattr(weight, "grad_ready_hook", None)
        if hook is not None and getattr(weight, "b200_wgrad_store_first", False):
            hook()
        return None
    dw = torch.zeros(weight.shape, dtype=torch.float32, device=weight.device)
    run(dw)
    return dw.to(weight.dtype)


class _GroupedLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xp, weight, lay: RaggedLayout, stride: int, fuse_wgrad: bool):
        ctx.lay, ctx.stride, ctx.fuse = lay, stride, fuse_wgrad
        ctx.save_for_backward(xp, weight)
        return grouped_gemm(xp, weight, lay, stride)

    @staticmethod
    def backward(ctx, dy):
        xp, weight = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = grouped_gemm(dy, weight, ctx.lay, ctx.stride, b_mn=True)
        if ctx.needs_input_grad[1]:
            dw = grouped_wgrad(dy, xp, weight, ctx.lay, ctx.stride, ctx.fuse)
        return dx, dw, None, None, None


def grouped_linear(xp, weight, lay: RaggedLayout, stride: int, fuse_wgrad: bool = False):
    return _GroupedLinear.apply(xp, weight, lay, stride, fuse_wgrad)


class _GroupedLinearCombine(torch.autograd.Function):
    """out[t] = sum_e p[t, e] * (h_rows(t, e) W2_e^T): expert FC2 with the router probability applied in the GEMM
    epilogue (per-row scale) and the un-permutation as a gather-sum; the backward produces d(probs) in the same pass
    that re-permutes and scales dOut."""

    @staticmethod
    def forward(ctx, h, weight, probs, lay: RaggedLayout, stride: int, fuse_wgrad: bool):
        ctx.lay, ctx.stride, ctx.fuse = lay, stride, fuse_wgrad
        z = grouped_gemm(h, weight, lay, stride, row_scale=lay.row_prob)
        out = _gather_sum(z, lay)
        ctx.save_for_backward(h, weight, z)
        ctx.probs_shape, ctx.probs_dtype = probs.shape, probs.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        h, weight, z = ctx.saved_tensors
        lay = ctx.lay
        dout = dout.contiguous()
        dprobs = torch.zeros(ctx.probs_shape, dtype=torch.float32, device=dout.device)
        dy = _gather_rows(dout, lay, lay.E_total, scale=lay.row_prob, other=z, dot_out=dprobs)
        dh = grouped_gemm(dy, weight, lay, ctx.stride, b_mn=True) if ctx.needs_input_grad[0] else None
        dw = grouped_wgrad(dy, h, weight, lay, ctx.stride, ctx.fuse) if ctx.needs_input_grad[1] else None
        return dh, dw, dprobs.to(ctx.probs_dtype), None, None, None


def grouped_linear_combine(h, weight, probs, lay: RaggedLayout, stride: int, fuse_wgrad: bool = False):
    return _GroupedLinearCombine.apply(h, weight, probs, lay, stride, fuse_wgrad)


# ---------------------------------------------------------------------------------------------- router
class _RouterTopK(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, k, pre_softmax, scaling):
        lf = logits.float().contiguous()
        probs, rmap, tpe = ext().moe_router_topk_fwd(lf, k, pre_softmax, scaling)
        ctx.save_for_backward(lf, probs, rmap)
        ctx.pre, ctx.scaling, ctx.in_dtype = pre_softmax, scaling, logits.dtype
        ctx.mark_non_differentiable(rmap, tpe)
        return probs.to(logits.dtype), rmap, tpe

    @staticmethod
    def backward(ctx, dprobs, _dm, _dt):
        lf, probs, rmap = ctx.saved_tensors
        dl = ext().moe_router_topk_bwd(lf, probs, rmap, dprobs.float().contiguous(), ctx.pre, ctx.scaling)
        return dl.to(ctx.in_dtype), None, None, None


def router_topk(logits: torch.Tensor, k: int, pre_softmax: bool = False, scaling: Optional[float] = None):
    """Fused softmax + top-k -> (dense probs [n, E], routing map [n, E] bool, tokens_per_expert [E] int32)."""
    return _RouterTopK.apply(logits, k, pre_softmax, float(scaling) if scaling else 1.0)


def grouped_gemm_is_available() -> bool:
    """The grouped GEMM is part of the framework's own extension (no external ``grouped_gemm`` package)."""
    from . import have_ext

    return have_ext()


def assert_grouped_gemm_is_available() -> None:
    if not grouped_gemm_is_available():
        raise AssertionError("the grouped tcgen05 GEMM needs the compiled extension (python -c 'import __graft_entry__ as g; g.build()')")

