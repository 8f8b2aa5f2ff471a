"""Operator layer: every hot op has (a) a hand-written sm_100a kernel reached through
``megatronapp_b200._C`` and (b) a plain PyTorch fp32 reference used on CPU (gloo plumbing
tests) and as the numerics oracle in ``tests/``.

CUDA tensors *never* silently fall back: if the extension is missing on a GPU box the
import raises.  The only library call on the GPU path is ``torch.matmul`` for operand
shapes/dtypes the tcgen05 GEMM does not take (fp32 params, K % 8 != 0), counted in
``fallback_counts`` so tests can assert the native path ran.
"""
from __future__ import annotations

import importlib.util
import math
import os
from collections import Counter
from pathlib import Path
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

_EXT = None
_EXT_ERR: Optional[BaseException] = None
fallback_counts: Counter = Counter()

EPI_STORE16, EPI_ACCUM_F32, EPI_STORE_F32 = 0, 1, 2


def _load_ext():
    global _EXT, _EXT_ERR
    if _EXT is not None or _EXT_ERR is not None:
        return _EXT
    so = Path(__file__).resolve().parent.parent / "_C.so"
    try:
        if not so.exists():
            from ..build import build_all

            build_all()
        spec = importlib.util.spec_from_file_location("megatronapp_b200._C", str(so))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _EXT = mod
    except BaseException as e:  # noqa: BLE001 - surfaced by ext() on the GPU path
        _EXT_ERR = e
    return _EXT


def ext():
    """The native module; raises if it cannot be loaded (never a silent fallback on GPU)."""
    m = _load_ext()
    if m is None:
        raise RuntimeError(f"megatronapp_b200._C (sm_100a extension) is not available: {_EXT_ERR!r}")
    return m


def have_ext() -> bool:
    return _load_ext() is not None


def launch_count() -> int:
    return ext().launch_count() if have_ext() else 0


def reset_launch_count() -> None:
    if have_ext():
        ext().reset_launch_count()


def _native(t: torch.Tensor) -> bool:
    return t.is_cuda and t.dtype in (torch.bfloat16, torch.float16)


# ---------------------------------------------------------------------------
# GEMM primitive:  D[M,N] = A[M,K] @ B[N,K]^T   (operands optionally stored transposed)
# ---------------------------------------------------------------------------

def gemm_ref(a, b, a_mn=False, b_mn=False, bias=None, act=0, out_dtype=None):
    A = a.t() if a_mn else a
    Bt = b if b_mn else b.t()
    d = A.float() @ Bt.float()
    if bias is not None:
        d = d + bias.float()
    if act == 1:
        d = F.gelu(d, approximate="tanh")
    return d.to(out_dtype or a.dtype)


def _gemm_ok(a, b, d) -> bool:
    if not (_native(a) and a.dtype == b.dtype):
        return False
    for t in (a, b, d):
        if t.dim() != 2 or t.stride(1) != 1 or t.stride(0) % 8 != 0 or t.data_ptr() % 16 != 0:
            return False
    return True


def gemm(a, b, out=None, *, a_mn=False, b_mn=False, epilogue=EPI_STORE16, bias=None, act=0,
         out_pre=None, block_n=0):
    """tcgen05 GEMM.  ``epilogue``: 0 store 16-bit, 1 ``out += acc`` (fp32), 2 ``out = acc`` (fp32)."""
    M = a.shape[1] if a_mn else a.shape[0]
    N = b.shape[1] if b_mn else b.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=a.dtype if epilogue == EPI_STORE16 else torch.float32,
                          device=a.device)
    if a.is_cuda and _gemm_ok(a, b, out) and (epilogue == EPI_STORE16 or out.shape[1] % 4 == 0):
        ext().gemm(a, b, out, a_mn, b_mn, epilogue, bias, act, out_pre, block_n)
        return out
    if a.is_cuda:
        fallback_counts["gemm"] += 1
    # reference / library path (CPU, fp32 params, odd shapes)
    A = a.t() if a_mn else a
    Bt = b if b_mn else b.t()
    if a.is_cuda:
        d = torch.matmul(A, Bt)
        if bias is not None:
            d = d + bias
        if out_pre is not None:
            out_pre.copy_(d)
        if act == 1:
            d = F.gelu(d.float(), approximate="tanh").to(d.dtype)
    else:
        d = A.float() @ Bt.float()
        if bias is not None:
            d = d + bias.float()
        if out_pre is not None:
            out_pre.copy_(d)
        if act == 1:
            d = F.gelu(d, approximate="tanh")
    if epilogue == EPI_ACCUM_F32:
        out.add_(d.to(out.dtype))
    else:
        out.copy_(d.to(out.dtype))
    return out


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b with tcgen05 forward / dgrad / wgrad.

    With ``weight.main_grad`` present and ``fuse_wgrad`` the wgrad GEMM accumulates straight into
    the fp32 DDP buffer (epilogue ``+=``) and notifies the bucket through
    ``weight.grad_ready_hook`` (capability of reference layers.py:509-541, without the
    dummy-gradient trick).
    """

    @staticmethod
    def forward(ctx, x, weight, bias, fuse_wgrad):
        ctx.fuse_wgrad = fuse_wgrad
        ctx.has_bias = bias is not None
        x2 = x.reshape(-1, x.shape[-1])
        out = gemm(x2, weight, bias=bias)
        ctx.save_for_backward(x2, weight)
        ctx.x_shape = x.shape
        return out.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm(dy2, weight, b_mn=True).view(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            dw = wgrad(dy2, x2, weight, ctx.fuse_wgrad)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = bias_grad(dy2).to(dy.dtype)
        return dx, dw, db, None


def wgrad(dy2, x2, weight, fuse: bool):
    """dW = dY^T X; accumulate into ``weight.main_grad`` when fused (returns None then)."""
    main_grad = getattr(weight, "main_grad", None)
    if fuse and main_grad is not None:
        first = not getattr(weight, "grad_added_to_main_grad", False) and getattr(weight, "b200_wgrad_store_first", False)
        if main_grad.dtype == torch.float32:
            gemm(dy2, x2, out=main_grad, a_mn=True, b_mn=True,
                 epilogue=EPI_STORE_F32 if first else EPI_ACCUM_F32)
        elif first:
            gemm(dy2, x2, out=main_grad, a_mn=True, b_mn=True)
        else:
            main_grad.add_(gemm(dy2, x2, a_mn=True, b_mn=True))
        weight.grad_added_to_main_grad = True
        # only weights owned by a linear layer have exactly one contribution per microbatch; a
        # tied embedding weight also gets the lookup gradient through autograd, whose
        # post-accumulate hook is what marks it ready
        if getattr(weight, "b200_wgrad_store_first", False):
            hook = getattr(weight, "grad_ready_hook", None)
            if hook is not None:
                hook()
        return None
    return gemm(dy2, x2, a_mn=True, b_mn=True)


def linear(x, weight, bias=None, fuse_wgrad: bool = False):
    return _LinearFn.apply(x, weight, bias, fuse_wgrad)


def bias_grad(dy2: torch.Tensor) -> torch.Tensor:
    """fp32 column sum of a [rows, cols] gradient."""
    if _native(dy2) and dy2.shape[-1] % 8 == 0:
        out = torch.empty(dy2.shape[-1], dtype=torch.float32, device=dy2.device)
        ext().bias_grad(dy2.contiguous(), out, False)
        return out
    return dy2.float().sum(dim=0)


# ---------------------------------------------------------------------------
# LayerNorm / RMSNorm
# ---------------------------------------------------------------------------

def norm_ref(x, gamma, beta, eps, rms=False, zero_centered=False):
    xf = x.float()
    g = gamma.float() + 1.0 if zero_centered else gamma.float()
    if rms:
        y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * g
    else:
        mu = xf.mean(-1, keepdim=True)
        var = (xf - mu).pow(2).mean(-1, keepdim=True)
        y = (xf - mu) * torch.rsqrt(var + eps) * g
        if beta is not None:
            y = y + beta.float()
    return y.to(x.dtype)


class _NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps, rms, zero_centered):
        xc = x.contiguous()
        y, mean, rstd = ext().layernorm_fwd(xc, gamma, beta, eps, rms, zero_centered)
        ctx.save_for_backward(xc, gamma, mean, rstd)
        ctx.rms, ctx.zc, ctx.has_beta = rms, zero_centered, beta is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        dx, dgamma, dbeta = ext().layernorm_bwd(dy.contiguous(), x, gamma, mean, rstd, ctx.rms, ctx.zc,
                                                ctx.has_beta)
        return dx, dgamma.to(gamma.dtype), (dbeta.to(gamma.dtype) if ctx.has_beta else None), None, None, None


def layer_norm(x, gamma, beta=None, eps=1e-5, rms=False, zero_centered=False):
    if _native(x) and gamma.dtype == x.dtype and x.shape[-1] % 8 == 0 and x.shape[-1] <= 8192:
        return _NormFn.apply(x, gamma, beta, eps, rms, zero_centered)
    if x.is_cuda:
        fallback_counts["norm"] += 1
    return norm_ref(x, gamma, beta, eps, rms, zero_centered)


# ---------------------------------------------------------------------------
# activations
# ---------------------------------------------------------------------------

def bias_gelu_ref(x, bias=None):
    xf = x.float() + (bias.float() if bias is not None else 0.0)
    return F.gelu(xf, approximate="tanh").to(x.dtype)


class _BiasGeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias):
        xc = x.contiguous()
        ctx.save_for_backward(xc, bias)
        return ext().bias_gelu_fwd(xc, bias)

    @staticmethod
    def backward(ctx, dy):
        x, bias = ctx.saved_tensors
        dx = ext().bias_gelu_bwd(dy.contiguous(), x, bias)
        db = None
        if bias is not None and ctx.needs_input_grad[1]:
            db = bias_grad(dx.view(-1, dx.shape[-1])).to(bias.dtype)
        return dx, db


def bias_gelu(x, bias=None):
    if _native(x) and x.shape[-1] % 8 == 0:
        return _BiasGeluFn.apply(x, bias)
    return bias_gelu_ref(x, bias)


def swiglu_ref(x, bias=None):
    xf = x.float() + (bias.float() if bias is not None else 0.0)
    g, u = xf.chunk(2, dim=-1)
    return (F.silu(g) * u).to(x.dtype)


class _SwigluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias):
        xc = x.contiguous()
        ctx.save_for_backward(xc, bias)
        return ext().swiglu_fwd(xc, bias)

    @staticmethod
    def backward(ctx, dy):
        x, bias = ctx.saved_tensors
        dx = ext().swiglu_bwd(dy.contiguous(), x, bias)
        db = None
        if bias is not None and ctx.needs_input_grad[1]:
            db = bias_grad(dx.view(-1, dx.shape[-1])).to(bias.dtype)
        return dx, db


def swiglu(x, bias=None):
    if _native(x) and x.shape[-1] % 16 == 0:
        return _SwigluFn.apply(x, bias)
    return swiglu_ref(x, bias)


# ---------------------------------------------------------------------------
# bias + dropout + residual add
# ---------------------------------------------------------------------------

def bias_dropout_add_ref(x, bias, residual, p, training):
    y = x if bias is None else x + bias
    y = F.dropout(y, p=p, training=training)
    return y if residual is None else residual + y


class _BdaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, residual, p, seed, offset):
        out, mask = ext().bias_dropout_add(x.contiguous(), bias, residual, p, seed, offset)
        ctx.p = p
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.save_for_backward(mask)
        return out

    @staticmethod
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        dx = dy if ctx.p == 0.0 else ext().dropout_bwd(dy, mask, ctx.p)
        db = bias_grad(dx.reshape(-1, dx.shape[-1])).to(dy.dtype) if (ctx.has_bias and ctx.needs_input_grad[1]) else None
        return dx, db, (dy if ctx.has_res else None), None, None, None


_BDA_OFFSET = [0]


def bias_dropout_add(x, bias, residual, p: float, training: bool):
    p_eff = p if training else 0.0
    if _native(x) and x.shape[-1] % 8 == 0 and (residual is None or residual.dtype == x.dtype):
        if p_eff > 0.0:
            seed = int(torch.cuda.initial_seed() & 0x7FFFFFFFFFFFFFFF)
            _BDA_OFFSET[0] += 1
            return _BdaFn.apply(x, bias, residual, p_eff, seed, _BDA_OFFSET[0])
        return _BdaFn.apply(x, bias, residual, 0.0, 0, 0)
    return bias_dropout_add_ref(x, bias, residual, p, training)


# ---------------------------------------------------------------------------
# RoPE.  x: [s, b, h, d]; freqs: [s, rot_dim/2] angles; cos_sin: [s, rot_dim] = cos | sin
# ---------------------------------------------------------------------------

def rope_cos_sin(seq_len: int, rot_dim: int, base: float = 10000.0, device=None, scaling: float = 1.0,
                 offset: int = 0) -> torch.Tensor:
    inv = 1.0 / (base ** (torch.arange(0, rot_dim, 2, dtype=torch.float32, device=device) / rot_dim))
    pos = (torch.arange(offset, offset + seq_len, dtype=torch.float32, device=device)) / scaling
    ang = torch.outer(pos, inv)
    return torch.cat([ang.cos(), ang.sin()], dim=-1).contiguous()


def rope_ref(x, cos_sin, interleaved=False, inverse=False):
    """``cos_sin``: [s, rot] (one table for the batch) or [s, b, rot] (per-sample positions: M-RoPE)."""
    s = x.shape[0]
    rot = cos_sin.shape[-1]
    half = rot // 2
    if cos_sin.dim() == 3:
        cos = cos_sin[:s, :, :half].float()[:, :, None, :]
        sin = cos_sin[:s, :, half:].float()[:, :, None, :]
    else:
        cos = cos_sin[:s, :half].float()[:, None, None, :]
        sin = cos_sin[:s, half:].float()[:, None, None, :]
    if inverse:
        sin = -sin
    xf = x.float()
    xr, xp = xf[..., :rot], xf[..., rot:]
    if interleaved:
        x0, x1 = xr[..., 0::2], xr[..., 1::2]
        y = torch.stack([x0 * cos - x1 * sin, x1 * cos + x0 * sin], dim=-1).flatten(-2)
    else:
        x0, x1 = xr[..., :half], xr[..., half:]
        y = torch.cat([x0 * cos - x1 * sin, x1 * cos + x0 * sin], dim=-1)
    return torch.cat([y, xp], dim=-1).to(x.dtype)


class _RopeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cos_sin, interleaved):
        ctx.save_for_backward(cos_sin)
        ctx.interleaved = interleaved
        return ext().rope_apply(x.contiguous(), cos_sin, interleaved, False)

    @staticmethod
    def backward(ctx, dy):
        (cos_sin,) = ctx.saved_tensors
        return ext().rope_apply(dy.contiguous(), cos_sin, ctx.interleaved, True), None, None


def rope(x, cos_sin, interleaved=False):
    if _native(x) and x.dim() == 4 and cos_sin.dim() == 2 and cos_sin.dtype == torch.float32 and cos_sin.is_contiguous():
        return _RopeFn.apply(x, cos_sin, interleaved)
    return rope_ref(x, cos_sin, interleaved)


# ---------------------------------------------------------------------------
# attention.  q,k,v: [b, h, s, d]
# ---------------------------------------------------------------------------

def attention_ref(q, k, v, scale=None, causal=True, dropout_p=0.0, return_probs=False, mask=None, bias=None):
    b, hq, sq, d = q.shape
    hkv = k.shape[1]
    if hkv != hq:
        k = k.repeat_interleave(hq // hkv, dim=1)
        v = v.repeat_interleave(hq // hkv, dim=1)
    scale = scale if scale is not None else 1.0 / math.sqrt(d)
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    if bias is not None:                    # additive [1|b, h, sq, sk] (T5 relative position bias)
        s = s + bias.float()
    sk = k.shape[2]
    if causal:
        cm = torch.ones(sq, sk, dtype=torch.bool, device=q.device).tril(diagonal=sk - sq)
        s = s.masked_fill(~cm, float("-inf"))
    if mask is not None:  # True = masked out (Megatron convention); finite fill keeps fully
        s = s.masked_fill(mask, -10000.0)  # masked (padding) query rows NaN-free, as in the reference
    p = torch.softmax(s, dim=-1)
    if dropout_p > 0.0:
        p = F.dropout(p, p=dropout_p)
    o = torch.matmul(p, v.float()).to(q.dtype)
    return (o, p) if return_probs else o


class _AttnFn(torch.autograd.Function):
    """tcgen05 flash attention on strided [b, h, s, d] views (csrc/attention_sm100.cu)."""

    @staticmethod
    def forward(ctx, q, k, v, scale, causal):
        o, lse, _ = ext().attn_fwd(q, k, v, scale, causal, False)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.scale, ctx.causal = scale, causal
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        dq, dk, dv = ext().attn_bwd(do, q, k, v, o, lse, ctx.scale, ctx.causal, None, None, None)
        return dq, dk, dv, None, None


class _PackedQKVAttnFn(torch.autograd.Function):
    """Self-attention straight on the packed projection output ``[s, b, np, 3*hn]`` (MHA): q / k / v are
    strided views of it, and the backward kernels write dq / dk / dv into the three slices of one packed
    gradient tensor -- no split copies forward, no concatenation backward."""

    @staticmethod
    def forward(ctx, mixed, scale, causal):
        hn = mixed.shape[-1] // 3
        q, k, v = (mixed[..., i * hn:(i + 1) * hn].permute(1, 2, 0, 3) for i in range(3))
        o, lse, _ = ext().attn_fwd(q, k, v, scale, causal, False)
        ctx.save_for_backward(mixed, o, lse)
        ctx.scale, ctx.causal = scale, causal
        return o

    @staticmethod
    def backward(ctx, do):
        mixed, o, lse = ctx.saved_tensors
        hn = mixed.shape[-1] // 3
        q, k, v = (mixed[..., i * hn:(i + 1) * hn].permute(1, 2, 0, 3) for i in range(3))
        dmixed = torch.empty_like(mixed)
        dq, dk, dv = (dmixed[..., i * hn:(i + 1) * hn].permute(1, 2, 0, 3) for i in range(3))
        ext().attn_bwd(do, q, k, v, o, lse, ctx.scale, ctx.causal, dq, dk, dv)
        return dmixed, None, None


def attention_packed_qkv(mixed, scale=None, causal=True):
    """``mixed``: ``[s, b, np, 3*hn]`` (q | k | v per head).  Returns ``[b, np, s, hn]`` or ``None`` when the
    native packed path does not apply (caller falls back to split + ``attention``)."""
    hn = mixed.shape[-1] // 3
    if not (_native(mixed) and _ATTN_IMPL in ("auto", "native") and hn in (64, 128) and mixed.is_contiguous()
            and mixed.shape[0] % 64 == 0 and native_attention_available()):
        return None
    scale = scale if scale is not None else 1.0 / math.sqrt(hn)
    return _PackedQKVAttnFn.apply(mixed, scale, causal)


_ATTN_IMPL = os.environ.get("B200_ATTN_IMPL", "auto")  # auto | native | sdpa | ref
_NATIVE_ATTN_OK = None


def native_attention_available() -> bool:
    """True when the in-tree tcgen05 attention kernels (forward AND backward) run on this device."""
    global _NATIVE_ATTN_OK
    if _NATIVE_ATTN_OK is None:
        _NATIVE_ATTN_OK = False
        if have_ext() and torch.cuda.is_available():
            try:
                q = torch.zeros(1, 1, 128, 64, dtype=torch.bfloat16, device="cuda")
                o, lse, _ = ext().attn_fwd(q, q, q, 1.0, True, False)
                ext().attn_bwd(o, q, q, q, o, lse, 1.0, True, None, None, None)
                _NATIVE_ATTN_OK = True
            except RuntimeError:
                _NATIVE_ATTN_OK = False
    return _NATIVE_ATTN_OK


def _attn_view_ok(t) -> bool:
    return (t.stride(3) == 1 and t.stride(0) % 8 == 0 and t.stride(1) % 8 == 0 and t.stride(2) % 8 == 0
            and t.data_ptr() % 16 == 0)


def attention(q, k, v, scale=None, causal=True, dropout_p=0.0, training=True, mask=None, bias=None):
    """Fused attention on logical ``[b, h, s, d]`` tensors (any strides with a contiguous last dim,
    e.g. permuted ``[s, b, h, d]`` activations -- consumed in place through 4-D tensor maps).
    Native tcgen05 kernels when usable; otherwise the SDPA library kernel on CUDA (counted in
    ``fallback_counts['attention_sdpa']``) or the reference on CPU."""
    d = q.shape[-1]
    scale = scale if scale is not None else 1.0 / math.sqrt(d)
    p = dropout_p if training else 0.0
    if not q.is_cuda or bias is not None:       # a learned additive bias needs d(bias): autograd through the reference
        return attention_ref(q, k, v, scale, causal, p, mask=mask, bias=bias)
    impl = _ATTN_IMPL
    sq, sk = q.shape[2], k.shape[2]
    needs_grad = torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad)
    usable = (_native(q) and p == 0.0 and mask is None and d in (64, 128) and (not causal or sq == sk)
              and q.shape[1] % k.shape[1] == 0 and (not needs_grad or (sq % 64 == 0 and sk % 64 == 0)))
    if impl in ("auto", "native") and usable and native_attention_available():
        q, k, v = (t if _attn_view_ok(t) else t.contiguous() for t in (q, k, v))
        return _AttnFn.apply(q, k, v, scale, causal)
    if impl == "native":
        raise RuntimeError("native attention requested but not usable for this problem")
    if impl == "ref":
        return attention_ref(q, k, v, scale, causal, p, mask=mask)
    fallback_counts["attention_sdpa"] += 1
    am = None if mask is None else torch.zeros(mask.shape, dtype=q.dtype, device=q.device).masked_fill(mask, -10000.0)
    return F.scaled_dot_product_attention(q, k, v, attn_mask=am, dropout_p=p,
                                          is_causal=causal and mask is None, scale=scale,
                                          enable_gqa=(k.shape[1] != q.shape[1]))


def attention_varlen(q, k, v, cu_seqlens_q, cu_seqlens_kv=None, scale=None, causal=True, dropout_p=0.0, training=True):
    """Packed (THD) attention: ``q/k/v [t, h, d]`` hold several sequences back to back, ``cu_seqlens`` their boundaries;
    tokens attend within their own sequence only.  Every sequence runs through ``attention`` (the tcgen05 flash
    kernel when its length qualifies), so no ``t x t`` mask is ever materialised.  Returns ``[t, h, d]``."""
    if isinstance(cu_seqlens_q, tuple) and len(cu_seqlens_q) == 2 and isinstance(cu_seqlens_q[0], list):
        bq, bk = cu_seqlens_q                        # host boundaries cached on PackedSeqParams: no device sync
    else:
        cu_kv = cu_seqlens_q if cu_seqlens_kv is None else cu_seqlens_kv
        to_list = lambda t: [int(x) for x in (t.tolist() if torch.is_tensor(t) else t)]  # noqa: E731
        bq, bk = to_list(cu_seqlens_q), to_list(cu_kv)
    outs = []
    for (q0, q1), (k0, k1) in zip(zip(bq[:-1], bq[1:]), zip(bk[:-1], bk[1:])):
        if q1 == q0:
            continue
        o = attention(q[q0:q1].transpose(0, 1).unsqueeze(0), k[k0:k1].transpose(0, 1).unsqueeze(0),
                      v[k0:k1].transpose(0, 1).unsqueeze(0), scale, causal, dropout_p, training)
        outs.append(o[0].transpose(0, 1))
    tail = q.shape[0] - bq[-1]                       # padding tokens after the last sequence produce zeros
    if tail > 0:
        outs.append(q.new_zeros(tail, q.shape[1], v.shape[-1]))
    return torch.cat(outs, dim=0)


# ---------------------------------------------------------------------------
# (vocab-parallel) cross entropy on [rows, vocab_local] logits
# ---------------------------------------------------------------------------

def cross_entropy_ref(logits, labels, label_smoothing=0.0):
    lf = logits.float()
    return F.cross_entropy(lf.view(-1, lf.shape[-1]), labels.view(-1), reduction="none",
                           label_smoothing=label_smoothing).view(labels.shape)


class _CrossEntropyFn(torch.autograd.Function):
    """Per-token loss with one packed TP all-reduce pair (max, then [sumexp | target]).

    Saves only the logits; the backward kernel recomputes softmax and may write the gradient
    in place of the logits (reference cross_entropy.py:123-220 keeps a full softmax copy)."""

    @staticmethod
    def forward(ctx, logits, labels, vocab_start, vocab_total, group, label_smoothing, inplace_grad):
        import torch.distributed as dist

        l2 = logits.view(-1, logits.shape[-1])
        lab = labels.reshape(-1).contiguous()
        mx, se, tg = ext().ce_local_stats(l2, lab, vocab_start)
        tp = group is not None and dist.is_initialized() and dist.get_world_size(group) > 1
        if tp:
            gmax = mx.clone()
            dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
            ext().ce_rescale(se, mx, gmax)
            packed = torch.stack([se, tg])
            dist.all_reduce(packed, group=group)
            se, tg = packed[0].contiguous(), packed[1].contiguous()
        else:
            gmax = mx
        loss = torch.log(se) + gmax - tg
        if label_smoothing > 0.0:
            # smoothed target needs the mean log-prob over the vocab
            mean_logit = l2.float().sum(-1)
            if tp:
                dist.all_reduce(mean_logit, group=group)
            mean_logp = mean_logit / vocab_total - (torch.log(se) + gmax)
            k = label_smoothing * vocab_total / (vocab_total - 1)
            loss = (1.0 - k) * loss - k * mean_logp
        ctx.save_for_backward(l2, lab, gmax, se)
        ctx.vocab_start, ctx.vocab_total = vocab_start, vocab_total
        ctx.ls, ctx.inplace = label_smoothing, inplace_grad
        ctx.shape = logits.shape
        return loss.view(labels.shape)

    @staticmethod
    def backward(ctx, dloss):
        l2, lab, gmax, se = ctx.saved_tensors
        dl = l2 if ctx.inplace else torch.empty_like(l2)
        g = dloss.reshape(-1).float().contiguous()
        ext().ce_grad(l2, lab, gmax, se, g, dl, ctx.vocab_start, ctx.ls, ctx.vocab_total)
        return dl.view(ctx.shape), None, None, None, None, None, None


def cross_entropy(logits, labels, vocab_start=0, vocab_total=None, group=None, label_smoothing=0.0,
                  inplace_grad=True):
    vocab_total = vocab_total or logits.shape[-1]
    if logits.is_cuda and logits.dtype in (torch.bfloat16, torch.float16, torch.float32) \
            and logits.stride(-1) == 1 and have_ext():
        lg = logits if logits.is_contiguous() else logits.contiguous()
        return _CrossEntropyFn.apply(lg, labels, vocab_start, vocab_total, group, label_smoothing,
                                     inplace_grad)
    from ..parallel.cross_entropy import vocab_parallel_cross_entropy_ref

    return vocab_parallel_cross_entropy_ref(logits, labels, vocab_start, vocab_total, group, label_smoothing)


# ---------------------------------------------------------------------------
# optimizer / multi-tensor
# ---------------------------------------------------------------------------

def adam_step_ref(master, exp_avg, exp_avg_sq, grad, param16, lr, beta1, beta2, eps, weight_decay,
                  grad_scale, step):
    g = grad.float() * grad_scale
    exp_avg.mul_(beta1).add_(g, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = exp_avg_sq.sqrt() / math.sqrt(bc2) + eps
    upd = (exp_avg / bc1) / denom + weight_decay * master
    master.add_(upd, alpha=-lr)
    if param16 is not None:
        param16.copy_(master)


def adam_step(master, exp_avg, exp_avg_sq, grad, param16, lr, beta1, beta2, eps, weight_decay,
              grad_scale, step, found_inf=None):
    if master.is_cuda and grad.dtype == torch.float32:
        ext().adam_step(master, exp_avg, exp_avg_sq, grad, param16, lr, beta1, beta2, eps, weight_decay,
                        grad_scale, step, found_inf)
    else:
        adam_step_ref(master, exp_avg, exp_avg_sq, grad, param16, lr, beta1, beta2, eps, weight_decay,
                      grad_scale, step)


def l2norm_sq(tensors, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Sum of squares over a list of tensors -> 1-element fp32 tensor (device side, no sync)."""
    tensors = [t for t in tensors if t is not None and t.numel() > 0]
    dev = tensors[0].device if tensors else "cpu"
    if out is None:
        out = torch.zeros(1, dtype=torch.float32, device=dev)
    for t in tensors:
        if t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.bfloat16, torch.float16):
            ext().l2norm_sq(t, out)
        else:
            out += t.float().pow(2).sum()
    return out


def scale_(t: torch.Tensor, scale: float = 1.0, scale_dev: Optional[torch.Tensor] = None):
    if t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.bfloat16, torch.float16):
        ext().scale_(t, scale, scale_dev)
    else:
        t.mul_(scale if scale_dev is None else scale * scale_dev.item())
    return t


def chunk_mean(x: torch.Tensor, pixels: int) -> torch.Tensor:
    """MegaScope compressor: mean over ``pixels`` contiguous chunks of the last dim."""
    if _native(x):
        return ext().chunk_mean(x, pixels)
    cols = x.shape[-1]
    outs = []
    for px in range(pixels):
        c0, c1 = px * cols // pixels, (px + 1) * cols // pixels
        outs.append(x[..., c0:max(c1, c0 + 1)].float().mean(-1))
    return torch.stack(outs, dim=-1)


# ---------------------------------------------------------------------------
# MegaScan software slow-rank injection: burn ``cycles`` SM clocks in front of every transformer layer,
# forward and backward, on the rank under test (what a clock-capped GPU looks like to its peers)
# ---------------------------------------------------------------------------
_SLOW_SPIN_CYCLES = 0


def set_slow_rank_spin(cycles: int) -> None:
    global _SLOW_SPIN_CYCLES
    _SLOW_SPIN_CYCLES = int(cycles)


class _SlowSpinFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if x.is_cuda:
            ext().spin(_SLOW_SPIN_CYCLES)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        if g.is_cuda:
            ext().spin(_SLOW_SPIN_CYCLES)
        return g


def slow_spin(x):
    return _SlowSpinFn.apply(x) if _SLOW_SPIN_CYCLES > 0 else x
