"""Numerics + performance checks of every sm_100a kernel against PyTorch fp32 references.

Each check runs in its own subprocess with a timeout so a hung / faulting kernel cannot take
the whole session (or the GPU lease) down; results are merged into one JSON file.

    python tools/gpu_checks.py --out gpurun_out/checks.json [--only gemm_nt,ln] [--perf]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _rel_err(a, b):
    """Element-wise error: max_i |a_i - b_i| / (|b_i| + rms(b)).  ``err <= tol`` is the element-wise criterion
    ``|a - b| <= tol * |b| + tol * rms(b)`` (rtol = tol, atol = tol * rms): an error on a small-magnitude output is
    measured against the typical magnitude of the tensor, not against its largest element."""
    import torch

    a, b = a.float(), b.float()
    if b.numel() == 0:
        return 0.0
    rms = b.pow(2).mean().sqrt().clamp_min(1e-12)
    return float(((a - b).abs() / (b.abs() + rms)).max())


def _time(fn, iters=20, warmup=5, flush=None):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


# ---------------------------------------------------------------------------------- checks

def check_gemm(mode, shapes=None, dtype="bf16"):
    import torch
    from megatronapp_b200 import ops

    dt = torch.bfloat16 if dtype == "bf16" else torch.float16
    res = {}
    shapes = shapes or [(128, 128, 64), (256, 256, 128), (384, 512, 320), (2048, 4096, 4096), (300, 520, 264),
                        (4096, 1024, 2048)]
    for (M, N, K) in shapes:
        if mode.startswith("t"):
            M = (M + 7) // 8 * 8   # A is stored [K, M]: TMA needs a 16-byte aligned row pitch
        for bn in (128, 256, 2):   # 2 = cta_group::2 kernel
            if bn == 2 and (M < 256 or N < 128):
                continue
            torch.manual_seed(M + N + K)
            A = torch.randn(M, K, device="cuda", dtype=dt) * 0.5
            B = torch.randn(N, K, device="cuda", dtype=dt) * 0.5
            bias = torch.randn(N, device="cuda", dtype=dt)
            ref = A.float() @ B.float().t()
            key = f"{mode}_{M}x{N}x{K}_bn{bn}"
            try:
                if mode == "nt":
                    out = ops.gemm(A, B, block_n=bn)
                    res[key] = _rel_err(out, ref)
                elif mode == "nt_bias_gelu":
                    pre = torch.empty(M, N, device="cuda", dtype=dt)
                    out = ops.gemm(A, B, bias=bias, act=1, out_pre=pre, block_n=bn)
                    r2 = ref + bias.float()
                    e1 = _rel_err(pre, r2)
                    e2 = _rel_err(out, torch.nn.functional.gelu(r2, approximate="tanh"))
                    res[key] = max(e1, e2)
                elif mode == "nn":   # B stored [K, N]
                    Bt = B.t().contiguous()
                    out = ops.gemm(A, Bt, b_mn=True, block_n=bn)
                    res[key] = _rel_err(out, ref)
                elif mode == "tn":   # A stored [K, M]
                    At = A.t().contiguous()
                    out = ops.gemm(At, B, a_mn=True, block_n=bn)
                    res[key] = _rel_err(out, ref)
                elif mode == "tt":   # both transposed: wgrad layout
                    At, Bt = A.t().contiguous(), B.t().contiguous()
                    out = ops.gemm(At, Bt, a_mn=True, b_mn=True, block_n=bn)
                    res[key] = _rel_err(out, ref)
                elif mode == "tt_accum":
                    At, Bt = A.t().contiguous(), B.t().contiguous()
                    if N % 4:
                        continue
                    acc = torch.randn(M, N, device="cuda", dtype=torch.float32)
                    want = acc + ref
                    ops.gemm(At, Bt, out=acc, a_mn=True, b_mn=True, epilogue=1, block_n=bn)
                    res[key] = _rel_err(acc, want)
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                res[key] = f"EXC {type(e).__name__}: {str(e)[:200]}"
    res["_fallbacks"] = dict(ops.fallback_counts)
    return res


def check_elementwise():
    import torch
    from megatronapp_b200 import ops

    res = {}
    dt = torch.bfloat16
    torch.manual_seed(0)
    for cols in (512, 4096, 5120):
        x = torch.randn(777, cols, device="cuda", dtype=dt, requires_grad=True)
        g = (torch.randn(cols, device="cuda", dtype=dt) * 0.1 + 1).requires_grad_(True)
        b = (torch.randn(cols, device="cuda", dtype=dt) * 0.1).requires_grad_(True)
        for rms in (False, True):
            y = ops.layer_norm(x, g, None if rms else b, 1e-5, rms=rms)
            dy = torch.randn_like(y)
            y.backward(dy)
            gx, gg = x.grad.clone(), g.grad.clone()
            gb = None if rms else b.grad.clone()
            x.grad = g.grad = None
            if not rms:
                b.grad = None
            xf, gf = x.detach().float().requires_grad_(True), g.detach().float().requires_grad_(True)
            bf = None if rms else b.detach().float().requires_grad_(True)
            yr = ops.norm_ref(xf, gf, bf, 1e-5, rms=rms)
            yr.backward(dy.float())
            tag = f"{'rms' if rms else 'ln'}_{cols}"
            res[tag + "_fwd"] = _rel_err(y, yr)
            res[tag + "_dx"] = _rel_err(gx, xf.grad)
            res[tag + "_dgamma"] = _rel_err(gg, gf.grad)
            if not rms:
                res[tag + "_dbeta"] = _rel_err(gb, bf.grad)
    # gelu / swiglu
    x = torch.randn(1000, 4096, device="cuda", dtype=dt, requires_grad=True)
    bias = torch.randn(4096, device="cuda", dtype=dt, requires_grad=True)
    y = ops.bias_gelu(x, bias)
    dy = torch.randn_like(y)
    y.backward(dy)
    xf, bf = x.detach().float().requires_grad_(True), bias.detach().float().requires_grad_(True)
    yr = ops.bias_gelu_ref(xf, bf)
    yr.backward(dy.float())
    res["gelu_fwd"], res["gelu_dx"], res["gelu_db"] = _rel_err(y, yr), _rel_err(x.grad, xf.grad), _rel_err(bias.grad, bf.grad)
    x.grad = bias.grad = None
    y = ops.swiglu(x, bias)
    dy = torch.randn_like(y)
    y.backward(dy)
    xf, bf = x.detach().float().requires_grad_(True), bias.detach().float().requires_grad_(True)
    yr = ops.swiglu_ref(xf, bf)
    yr.backward(dy.float())
    res["swiglu_fwd"], res["swiglu_dx"], res["swiglu_db"] = _rel_err(y, yr), _rel_err(x.grad, xf.grad), _rel_err(bias.grad, bf.grad)
    # bias dropout add (p = 0 exact; p > 0 statistical)
    h = torch.randn(64, 8, 1024, device="cuda", dtype=dt)
    r = torch.randn_like(h)
    b2 = torch.randn(1024, device="cuda", dtype=dt)
    o = ops.bias_dropout_add(h, b2, r, 0.0, True)
    res["bda_p0"] = _rel_err(o, (h.float() + b2.float() + r.float()))
    o = ops.bias_dropout_add(h, None, None, 0.25, True)
    kept = (o != 0).float().mean().item()
    res["bda_keep_frac(expect .75)"] = kept
    res["bda_scale"] = _rel_err(o[o != 0], (h.float() / 0.75)[o != 0])
    # rope
    q = torch.randn(128, 2, 8, 128, device="cuda", dtype=dt, requires_grad=True)
    cs = ops.rope_cos_sin(128, 128, device="cuda")
    for inter in (False, True):
        y = ops.rope(q, cs, inter)
        dy = torch.randn_like(y)
        y.backward(dy)
        qf = q.detach().float().requires_grad_(True)
        yr = ops.rope_ref(qf, cs, inter)
        yr.backward(dy.float())
        res[f"rope_fwd_i{int(inter)}"], res[f"rope_bwd_i{int(inter)}"] = _rel_err(y, yr), _rel_err(q.grad, qf.grad)
        q.grad = None
    cs2 = ops.rope_cos_sin(128, 64, device="cuda")
    y = ops.rope(q.detach(), cs2, False)
    res["rope_partial"] = _rel_err(y, ops.rope_ref(q.detach().float(), cs2, False))
    # bias grad
    d = torch.randn(3000, 4096, device="cuda", dtype=dt)
    res["bias_grad"] = _rel_err(ops.bias_grad(d), d.float().sum(0))
    # cross entropy
    for ldt in (torch.bfloat16, torch.float32):
        lg = (torch.randn(512, 50304, device="cuda", dtype=ldt) * 2).requires_grad_(True)
        lab = torch.randint(0, 50304, (512,), device="cuda")
        loss = ops.cross_entropy(lg.clone() if True else lg, lab, inplace_grad=False)
        lgf = lg.detach().float().requires_grad_(True)
        ref = ops.cross_entropy_ref(lgf, lab)
        res[f"ce_fwd_{str(ldt)[6:]}"] = _rel_err(loss, ref)
        w = torch.rand(512, device="cuda")
        (loss * w).sum().backward()
        (ref * w).sum().backward()
        res[f"ce_bwd_{str(ldt)[6:]}"] = _rel_err(lg.grad, lgf.grad)
    # adam
    n = 1_000_003
    m = torch.randn(n, device="cuda")
    ea, es = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    g = torch.randn(n, device="cuda")
    p16 = torch.empty(n, device="cuda", dtype=dt)
    m2, ea2, es2 = m.clone(), ea.clone(), es.clone()
    for step in (1, 2, 3):
        ops.adam_step(m, ea, es, g, p16, 1e-2, 0.9, 0.95, 1e-8, 0.1, 0.5, step)
        ops.adam_step_ref(m2, ea2, es2, g, None, 1e-2, 0.9, 0.95, 1e-8, 0.1, 0.5, step)
    res["adam_master"], res["adam_p16"] = _rel_err(m, m2), _rel_err(p16, m2)
    res["l2norm"] = abs(float(ops.l2norm_sq([g, m]).item()) / float((g.float().pow(2).sum() + m.pow(2).sum()).item()) - 1)
    res["chunk_mean"] = _rel_err(ops.chunk_mean(h, 64), ops.chunk_mean(h.float().cpu(), 64).cuda())
    res["_fallbacks"] = dict(ops.fallback_counts)
    return res


def check_gemm_perf():
    import torch
    from megatronapp_b200 import ops

    res = {}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    shapes = {"sq8192": (8192, 8192, 8192), "qkv_tp1": (2048, 12288, 4096), "fc1_tp1": (2048, 16384, 4096),
              "fc2_tp1": (2048, 4096, 16384), "proj_tp1": (2048, 4096, 4096), "fc1_mbs4": (8192, 16384, 4096),
              "lmhead": (2048, 50304, 4096), "qkv_tp4": (2048, 3072, 4096)}
    for name, (M, N, K) in shapes.items():
        A = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        B = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        fl = 2.0 * M * N * K
        for bn in (0, 256, 2):
            try:
                t = _time(lambda: ops.gemm(A, B, out=out, block_n=bn), flush=flush)
                res[f"{name}_b200_bn{bn}_tflops"] = round(fl / t / 1e9, 1)
            except Exception as e:  # noqa: BLE001
                res[f"{name}_b200_bn{bn}"] = f"EXC {e}"
        try:
            dYp = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
            dXp = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
            mgp = torch.zeros(N, K, device="cuda", dtype=torch.float32)
            for bn in (256, 2):
                t = _time(lambda: ops.gemm(dYp, B, out=dXp, b_mn=True, block_n=bn), flush=flush)
                res[f"{name}_dgrad_bn{bn}_tflops"] = round(fl / t / 1e9, 1)
                t = _time(lambda: ops.gemm(dYp, A, out=mgp, a_mn=True, b_mn=True, epilogue=1, block_n=bn), flush=flush)
                res[f"{name}_wgrad_bn{bn}_tflops"] = round(fl / t / 1e9, 1)
            del dYp, dXp, mgp
        except Exception as e:  # noqa: BLE001
            res[f"{name}_bwd_2cta"] = f"EXC {e}"
        t = _time(lambda: torch.matmul(A, B.t(), out=out), flush=flush)
        res[f"{name}_cublas_tflops"] = round(fl / t / 1e9, 1)
        # dgrad / wgrad layouts
        dY = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
        dX = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
        t = _time(lambda: ops.gemm(dY, B, out=dX, b_mn=True), flush=flush)
        res[f"{name}_dgrad_tflops"] = round(fl / t / 1e9, 1)
        mg = torch.zeros(N, K, device="cuda", dtype=torch.float32)
        t = _time(lambda: ops.gemm(dY, A, out=mg, a_mn=True, b_mn=True, epilogue=1), flush=flush)
        res[f"{name}_wgrad_accum_tflops"] = round(fl / t / 1e9, 1)
        del A, B, out, dY, dX, mg
    return res


def check_elementwise_perf():
    import torch
    from megatronapp_b200 import ops

    res = {}
    x = torch.randn(16384, 4096, device="cuda", dtype=torch.bfloat16)
    g = torch.ones(4096, device="cuda", dtype=torch.bfloat16)
    b = torch.zeros(4096, device="cuda", dtype=torch.bfloat16)
    nb = x.numel() * 2
    t = _time(lambda: ops.ext().layernorm_fwd(x, g, b, 1e-5, False, False))
    res["ln_fwd_GBs"] = round(2 * nb / t / 1e6, 1)
    y, mean, rstd = ops.ext().layernorm_fwd(x, g, b, 1e-5, False, False)
    t = _time(lambda: ops.ext().layernorm_bwd(x, x, g, mean, rstd, False, False, True))
    res["ln_bwd_GBs"] = round(3 * nb / t / 1e6, 1)
    t = _time(lambda: ops.ext().bias_gelu_fwd(x, b))
    res["gelu_fwd_GBs"] = round(2 * nb / t / 1e6, 1)
    n = 256 * 1024 * 1024
    m, ea, es, gr = (torch.zeros(n, device="cuda") for _ in range(4))
    p16 = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    t = _time(lambda: ops.adam_step(m, ea, es, gr, p16, 1e-3, 0.9, 0.95, 1e-8, 0.1, 1.0, 1), iters=5)
    res["adam_GBs"] = round(n * 30 / t / 1e6, 1)
    lg = torch.randn(2048, 50304, device="cuda", dtype=torch.bfloat16)
    lab = torch.randint(0, 50304, (2048,), device="cuda")
    t = _time(lambda: ops.ext().ce_local_stats(lg, lab, 0))
    res["ce_stats_GBs"] = round(lg.numel() * 2 / t / 1e6, 1)
    return res


_ATTN_CASES = [(1, 2, 2, 256, 128, True), (2, 4, 4, 512, 128, True), (2, 4, 2, 384, 64, True),
               (1, 2, 2, 256, 128, False), (1, 8, 8, 2048, 128, True), (1, 2, 1, 1024, 64, False)]


def _attn_inputs(b, h, hk, s, d, sbhd):
    """q/k/v as logical [b,h,s,d]; ``sbhd`` makes them permuted views of [s,b,h,d] storage (the
    layout the transformer layer hands over)."""
    import torch

    torch.manual_seed(s + d)
    if sbhd:
        mk = lambda hh: torch.randn(s, b, hh, d, device="cuda", dtype=torch.bfloat16).permute(1, 2, 0, 3)
    else:
        mk = lambda hh: torch.randn(b, hh, s, d, device="cuda", dtype=torch.bfloat16)
    return mk(h), mk(hk), mk(hk)


def check_attention_fwd():
    """Native tcgen05 flash-attention forward vs the fp32 reference (incl. ragged lengths)."""
    import torch
    from megatronapp_b200 import ops

    res = {}
    cases = [(c, True) for c in _ATTN_CASES] + [((1, 2, 2, 200, 128, False), False), ((1, 2, 2, 320, 64, True), False),
                                               ((1, 1, 1, 128, 128, True), False)]
    for (b, h, hk, s, d, causal), sbhd in cases:
        q, k, v = _attn_inputs(b, h, hk, s, d, sbhd)
        tag = f"attn_b{b}h{h}kv{hk}s{s}d{d}c{int(causal)}"
        try:
            o, lse, _ = ops.ext().attn_fwd(q, k, v, 1.0 / d ** 0.5, causal, False)
            torch.cuda.synchronize()
            orf = ops.attention_ref(q.float(), k.float(), v.float(), None, causal)
            kk = k.float().repeat_interleave(h // hk, dim=1)
            sc = torch.matmul(q.float(), kk.transpose(-1, -2)) / d ** 0.5
            if causal:
                sc = sc.masked_fill(~torch.ones(s, s, dtype=torch.bool, device="cuda").tril(), float("-inf"))
            res[tag + "_fwd"] = _rel_err(o, orf)
            res[tag + "_lse"] = _rel_err(lse, torch.logsumexp(sc, dim=-1))
        except Exception as e:  # noqa: BLE001
            res[tag] = f"EXC {type(e).__name__}: {str(e)[:200]}"
    return res


def check_attention():
    """Forward + backward (dq, dk, dv) of the native kernels through autograd."""
    import torch
    from megatronapp_b200 import ops

    res = {}
    for (b, h, hk, s, d, causal) in _ATTN_CASES:
        q, k, v = (t.detach().requires_grad_(True) for t in _attn_inputs(b, h, hk, s, d, True))
        tag = f"attn_b{b}h{h}kv{hk}s{s}d{d}c{int(causal)}"
        try:
            o = ops._AttnFn.apply(q, k, v, 1.0 / d ** 0.5, causal)
            do = torch.randn_like(o)
            o.backward(do)
            torch.cuda.synchronize()
            qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
            orf = ops.attention_ref(qf, kf, vf, None, causal)
            orf.backward(do.float())
            res[tag + "_fwd"] = _rel_err(o, orf)
            res[tag + "_dq"], res[tag + "_dk"], res[tag + "_dv"] = _rel_err(q.grad, qf.grad), _rel_err(k.grad, kf.grad), _rel_err(v.grad, vf.grad)
        except Exception as e:  # noqa: BLE001
            res[tag] = f"EXC {type(e).__name__}: {str(e)[:200]}"
    return res


def check_attention_perf():
    import torch
    from megatronapp_b200 import ops
    import torch.nn.functional as F

    res = {}
    b, h, s, d = 1, 32, 2048, 128
    # [s, b, h, d] storage viewed as [b, h, s, d]: what the transformer layer passes
    q, k, v = (torch.randn(s, b, h, d, device="cuda", dtype=torch.bfloat16).permute(1, 2, 0, 3).requires_grad_(True)
               for _ in range(3))
    fl_f = 4.0 * b * h * s * s * d / 2
    t = _time(lambda: F.scaled_dot_product_attention(q, k, v, is_causal=True))
    res["sdpa_fwd_tflops"] = round(fl_f / t / 1e9, 1)
    o = F.scaled_dot_product_attention(q, k, v, is_causal=True)
    do = torch.randn_like(o)
    t = _time(lambda: torch.autograd.grad(o, (q, k, v), do, retain_graph=True))
    res["sdpa_bwd_tflops"] = round(2.5 * fl_f / t / 1e9, 1)
    qd, kd, vd = q.detach(), k.detach(), v.detach()
    t = _time(lambda: ops.ext().attn_fwd(qd, kd, vd, 1.0 / d ** 0.5, True, False))
    res["b200_fwd_ms"], res["b200_fwd_tflops"] = round(t, 4), round(fl_f / t / 1e9, 1)
    o2, lse, _ = ops.ext().attn_fwd(qd, kd, vd, 1.0 / d ** 0.5, True, False)
    t = _time(lambda: ops.ext().attn_bwd(do, qd, kd, vd, o2, lse, 1.0 / d ** 0.5, True, None, None, None))
    res["b200_bwd_ms"], res["b200_bwd_tflops"] = round(t, 4), round(2.5 * fl_f / t / 1e9, 1)
    for (bb, hh) in [(4, 32), (4, 8)]:
        q4, k4, v4 = (torch.randn(s, bb, hh, d, device="cuda", dtype=torch.bfloat16).permute(1, 2, 0, 3) for _ in range(3))
        fl = 4.0 * bb * hh * s * s * d / 2
        t = _time(lambda: ops.ext().attn_fwd(q4, k4, v4, 1.0 / d ** 0.5, True, False))
        res[f"b200_fwd_tflops_b{bb}h{hh}"] = round(fl / t / 1e9, 1)
        t = _time(lambda: F.scaled_dot_product_attention(q4, k4, v4, is_causal=True))
        res[f"sdpa_fwd_tflops_b{bb}h{hh}"] = round(fl / t / 1e9, 1)
    return res


def check_fp8():
    """fp8 quantise kernels and the kind::f8f6f4 GEMM vs the dequantised fp32 product."""
    import torch
    from megatronapp_b200 import ops
    from megatronapp_b200.ops import fp8

    res = {}
    for (M, N, K) in [(256, 256, 128), (2048, 4096, 4096), (384, 520, 272), (4096, 1024, 2048)]:
        torch.manual_seed(M + K)
        a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.05
        bias = torch.randn(N, device="cuda", dtype=torch.bfloat16)
        key = f"{M}x{N}x{K}"
        try:
            a8, a8t, sa, am = fp8.quantize(a, False, True)
            b8, b8t, sb, _ = fp8.quantize(b, False, True)
            res[key + "_amax"] = abs(float(am) - float(a.float().abs().max())) / float(a.float().abs().max())
            ref_q = a.float() * (1.0 / sa)
            res[key + "_cast"] = _rel_err(fp8.dequantize(a8, sa), ref_q.clamp(-448, 448).to(torch.float8_e4m3fn).float() * sa)
            res[key + "_transpose"] = float((a8t != a8.t()).float().mean())
            ref = fp8.dequantize(a8, sa) @ fp8.dequantize(b8, sb).t()
            out = fp8.gemm_fp8(a8, b8, sa, sb, bias=bias)
            res[key + "_gemm"] = _rel_err(out, ref + bias.float())
            acc = torch.randn(M, N, device="cuda", dtype=torch.float32) if N % 4 == 0 else None
            if acc is not None:
                want = acc + ref
                fp8.gemm_fp8(a8, b8, sa, sb, out=acc, accumulate=True)
                res[key + "_accum"] = _rel_err(acc, want)
            d = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
            d8, d8t, ds, _ = fp8.quantize(d, True, True)       # E5M2 gradient operand
            refd = fp8.dequantize(d8, ds, True) @ fp8.dequantize(b8t, sb).t()
            res[key + "_dgrad_e5m2"] = _rel_err(fp8.gemm_fp8(d8, b8t, ds, sb, a_e5m2=True), refd)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            res[key] = f"EXC {type(e).__name__}: {str(e)[:200]}"
    # end-to-end linear vs bf16 reference (quantisation-level error)
    x = torch.randn(2048, 4096, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    w = (torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16) * 0.02).requires_grad_(True)
    y = fp8.fp8_linear(x, w)
    dy = torch.randn_like(y)
    y.backward(dy)
    ref = x.detach().float() @ w.detach().float().t()
    res["linear_fwd_vs_bf16"] = _rel_err(y, ref)
    res["linear_dx_vs_bf16"] = _rel_err(x.grad, dy.float() @ w.detach().float())
    res["linear_dw_vs_bf16"] = _rel_err(w.grad, dy.float().t() @ x.detach().float())
    res["_fallbacks"] = {k: v for k, v in ops.fallback_counts.items() if "fp8" in k}
    return res


def check_mxfp8():
    """Native block-scaled GEMM (tcgen05 kind::mxf8f6f4.block_scale, scales in TMEM) and the dual quantiser vs the
    PyTorch OCP-MX emulation; the mxfp8 linear (fprop / dgrad / wgrad) vs fp32."""
    import torch
    from megatronapp_b200 import ops
    from megatronapp_b200.ops import fp8

    res = {}
    for (M, N, K) in ((256, 384, 256), (384, 192, 128), (2048, 4096, 4096), (300, 520, 384), (4096, 1024, 2048)):
        torch.manual_seed(M + N + K)
        a = (torch.randn(M, K, device="cuda") * (torch.rand(M, 1, device="cuda") * 4 + 0.1)).bfloat16()
        b = (torch.randn(N, K, device="cuda") * 0.3).bfloat16()
        key = f"{M}x{N}x{K}"
        try:
            for e5 in (False, True):
                ar, _ = fp8.quantize_mx_native(a, e5, 0, -1)
                q_ref, s_ref = fp8.quantize_mx(a, e5)
                tag = key + ("_e5m2" if e5 else "_e4m3")
                res[tag + "_q_mismatch"] = float((ar.q != q_ref).float().mean())
                res[tag + "_sf_mismatch"] = float((fp8.sf_matrix(ar) != s_ref).float().mean())
            ar, _ = fp8.quantize_mx_native(a, False, 0, -1)
            br, _ = fp8.quantize_mx_native(b, False, 1, -1)
            res[key + "_sfb_mismatch"] = float((fp8.sf_matrix(br) != fp8.quantize_mx(b)[1]).float().mean())
            ref = fp8.dequantize_mx_native(ar) @ fp8.dequantize_mx_native(br).t()
            bias = torch.randn(N, device="cuda").bfloat16()
            out = fp8.gemm_mx_native(ar, br, bias=bias)
            res[key + "_gemm"] = _rel_err(out, ref + bias.float())
            acc = torch.randn(M, N, device="cuda")
            want = acc + ref
            fp8.gemm_mx_native(ar, br, out=acc, accumulate=True)
            res[key + "_accum"] = _rel_err(acc, want)
            if M % 128 == 0:
                _, ac = fp8.quantize_mx_native(a, True, -1, 0)            # a^T [K, M], blocks along M
                qt_ref, st_ref = fp8.quantize_mx(a.t().contiguous(), True)
                res[key + "_colwise_q_mismatch"] = float((ac.q != qt_ref).float().mean())
                res[key + "_colwise_sf_mismatch"] = float((fp8.sf_matrix(ac) != st_ref).float().mean())
        except Exception as e:  # noqa: BLE001
            res[key] = f"EXC {type(e).__name__}: {str(e)[:300]}"
    # the linear layer
    torch.manual_seed(0)
    x = (torch.randn(512, 1024, device="cuda") * 0.5).bfloat16().requires_grad_(True)
    w = (torch.randn(768, 1024, device="cuda") * 0.05).bfloat16().requires_grad_(True)
    rec = fp8.Fp8Recipe(fmt="hybrid", recipe="mxfp8")
    y = fp8.fp8_linear(x, w, None, rec, fp8._MxWeightCache())
    dy = torch.randn_like(y)
    y.backward(dy)
    res["linear_fwd_vs_fp32"] = _rel_err(y, x.detach().float() @ w.detach().float().t())
    res["linear_dx_vs_fp32"] = _rel_err(x.grad, dy.float() @ w.detach().float())
    res["linear_dw_vs_fp32"] = _rel_err(w.grad, dy.float().t() @ x.detach().float())
    res["_fallbacks"] = dict(ops.fallback_counts)
    return res


def check_mxfp8_perf():
    import torch
    from megatronapp_b200.ops import fp8

    res = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for (M, N, K, name) in ((8192, 16384, 4096, "fc1"), (8192, 4096, 16384, "fc2"), (8192, 8192, 8192, "sq8k")):
        a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        b = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
        ar, _ = fp8.quantize_mx_native(a, False, 0, -1)
        br, _ = fp8.quantize_mx_native(b, False, 1, -1)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        t = _time(lambda: fp8.gemm_mx_native(ar, br, out=out), flush=flush)
        res[f"mxfp8_{name}_tflops"] = 2.0 * M * N * K / t / 1e9
        t = _time(lambda: fp8.quantize_mx_native(a, False, 0, 1), flush=flush)
        res[f"mx_dual_quantize_{name}_GBs"] = (M * K * 2 + 2 * M * K) / t / 1e6
    return res


def check_fp8_perf():
    import torch
    from megatronapp_b200 import ops
    from megatronapp_b200.ops import fp8

    res = {}
    for name, (M, N, K) in {"sq8192": (8192, 8192, 8192), "fc1": (2048, 16384, 4096), "fc1_mbs4": (8192, 16384, 4096)}.items():
        a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
        a8, _, sa, _ = fp8.quantize(a)
        b8, _, sb, _ = fp8.quantize(b)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        t = _time(lambda: fp8.gemm_fp8(a8, b8, sa, sb, out=out))
        res[name + "_fp8_tflops"] = round(2.0 * M * N * K / t / 1e9, 1)
        t = _time(lambda: ops.gemm(a, b, out=out))
        res[name + "_bf16_tflops"] = round(2.0 * M * N * K / t / 1e9, 1)
        t = _time(lambda: fp8.quantize(a, False, True))
        res[name + "_quantize_GBs"] = round(a.numel() * 2 * 2 / t / 1e6, 1)    # amax read + cast read
    return res


def _ragged_layout(counts, n_tokens_hint=None):
    """A RaggedLayout with the given per-expert row counts (rows filled in order, like the permute kernel)."""
    import torch
    from megatronapp_b200.ops.moe import PAD, RaggedLayout

    E = len(counts)
    padded = [(c + PAD - 1) // PAD * PAD for c in counts]
    offs = [0]
    for p in padded:
        offs.append(offs[-1] + p)
    cap = offs[-1] + PAD          # some unused capacity behind the last segment
    row_slot = torch.full((cap,), -1, dtype=torch.int32)
    for e, c in enumerate(counts):
        row_slot[offs[e]:offs[e] + c] = torch.arange(c, dtype=torch.int32) * E + e
    dev = "cuda"
    return RaggedLayout(max(counts) if counts else 0, E, 0, E, cap, torch.tensor(counts, dtype=torch.int32, device=dev),
                        torch.tensor(offs, dtype=torch.int32, device=dev), row_slot.to(dev),
                        (row_slot >= 0).float().to(dev), torch.zeros(1, E, dtype=torch.int32, device=dev))


def check_grouped_gemm():
    """Grouped tcgen05 GEMM (forward, dgrad, wgrad) over the ragged 256-row padded layout vs per-group fp32 matmuls."""
    import torch
    from megatronapp_b200 import ops
    from megatronapp_b200.ops import moe as mops

    res = {}
    dt = torch.bfloat16
    for name, counts, H, F in (("e4", [300, 0, 1000, 257], 512, 768), ("e8", [700, 512, 90, 0, 2048, 1, 255, 256], 1024, 1536),
                              ("e2_big", [5000, 3100], 4096, 3584)):
        torch.manual_seed(len(counts) + H)
        lay = _ragged_layout(counts)
        E, cap = len(counts), lay.cap
        offs = lay.offsets.tolist()
        valid = (lay.row_slot >= 0)
        x = torch.randn(cap, H, device="cuda", dtype=dt) * 0.5
        x = x * valid.unsqueeze(-1)                               # padding rows are zero, as the permute kernel leaves them
        w = torch.randn(E * F, H, device="cuda", dtype=dt) * 0.05
        scale = torch.rand(cap, device="cuda") + 0.5
        y = mops.grouped_gemm(x, w, lay, F)
        ys = mops.grouped_gemm(x, w, lay, F, row_scale=scale)
        dy = torch.randn(cap, F, device="cuda", dtype=dt) * 0.5 * valid.unsqueeze(-1)
        dx = mops.grouped_gemm(dy, w, lay, F, b_mn=True)
        dw = torch.zeros(E * F, H, device="cuda", dtype=torch.float32)
        ops.ext().gemm_grouped(dy, x, dw, lay.offsets, 2, True, None, 0, None, None, E, F)
        e_y = e_ys = e_dx = e_dw = 0.0
        for g in range(E):
            r0, r1 = offs[g], offs[g + 1]
            wg = w[g * F:(g + 1) * F].float()
            ref_dw = dy[r0:r1].float().t() @ x[r0:r1].float()
            e_dw = max(e_dw, _rel_err(dw[g * F:(g + 1) * F], ref_dw) if r1 > r0 else float(dw[g * F:(g + 1) * F].abs().max()))
            if r1 == r0:
                continue
            ref = x[r0:r1].float() @ wg.t()
            e_y = max(e_y, _rel_err(y[r0:r1], ref))
            e_ys = max(e_ys, _rel_err(ys[r0:r1], ref * scale[r0:r1].unsqueeze(-1)))
            e_dx = max(e_dx, _rel_err(dx[r0:r1], dy[r0:r1].float() @ wg))
        res[f"{name}_fwd"], res[f"{name}_fwd_rowscale"], res[f"{name}_dgrad"], res[f"{name}_wgrad"] = e_y, e_ys, e_dx, e_dw
    res["_fallbacks"] = dict(ops.fallback_counts)
    return res


def check_moe_kernels():
    """Router top-k, layout builder, permute / combine kernels vs their PyTorch definitions."""
    import torch
    from megatronapp_b200 import ops
    from megatronapp_b200.moe.moe_utils import topk_softmax_with_capacity
    from megatronapp_b200.ops import moe as mops

    res = {}
    torch.manual_seed(3)
    for (n, E, k, pre) in ((1000, 8, 2, False), (777, 64, 6, True), (4096, 8, 2, True), (513, 128, 8, False)):
        logits = torch.randn(n, E, device="cuda") * 2
        l1 = logits.clone().requires_grad_(True)
        l2 = logits.clone().requires_grad_(True)
        p1, m1, t1 = mops.router_topk(l1, k, pre, 1.5)
        p2, m2, t2 = topk_softmax_with_capacity(l2, k, None, False, "probs", pre, None, None, 1.5)
        g = torch.randn_like(p1)
        (p1 * g).sum().backward()
        (p2 * g).sum().backward()
        tag = f"router_n{n}_E{E}_k{k}_pre{int(pre)}"
        res[tag + "_probs"] = float((p1 - p2).abs().max())
        res[tag + "_map_mismatch"] = float((m1 != m2).sum())
        res[tag + "_tpe"] = float((t1.long() - t2.long()).abs().max())
        res[tag + "_dlogits"] = float((l1.grad - l2.grad).abs().max())
        # layout: CUDA kernels vs the reference implementation of the same layout
        lay = mops.build_layout(m1.contiguous(), p1, 0, E, k)
        ref = mops.build_layout(m1.cpu(), p1.detach().cpu(), 0, E, k)
        res[tag + "_offsets"] = float((lay.offsets.cpu() - ref.offsets).abs().max())
        res[tag + "_counts"] = float((lay.counts.cpu() - ref.counts).abs().max())
        res[tag + "_dest_row"] = float((lay.dest_row.cpu() - ref.dest_row).abs().max())
        res[tag + "_row_slot"] = float((lay.row_slot.cpu() - ref.row_slot).abs().max())
        res[tag + "_row_prob"] = float((lay.row_prob.cpu() - ref.row_prob).abs().max())
        H = 256
        x = (torch.randn(n, H, device="cuda") * 0.5).bfloat16()
        xp = mops._gather_rows(x, lay, E)
        xr = mops._gather_rows(x.cpu(), ref, E)
        res[tag + "_permute"] = float((xp.cpu().float() - xr.float())[:int(ref.offsets[-1])].abs().max())
        y = (torch.randn(lay.cap, H, device="cuda") * 0.5).bfloat16()
        res[tag + "_gather_sum"] = _rel_err(mops._gather_sum(y, lay).cpu(), mops._gather_sum(y.cpu().float(), ref))
        dpr = torch.zeros(n, E, device="cuda")
        dpc = torch.zeros(n, E)
        dy = mops._gather_rows(x, lay, E, scale=lay.row_prob, other=y, dot_out=dpr)
        dyr = mops._gather_rows(x.cpu().float(), ref, E, scale=ref.row_prob, other=y.cpu().float(), dot_out=dpc)
        total = int(ref.offsets[-1])               # rows behind the last segment are never written (nor read)
        res[tag + "_combine_bwd_rows"] = _rel_err(dy.cpu()[:total], dyr[:total])
        res[tag + "_combine_bwd_dprob"] = _rel_err(dpr.cpu(), dpc)
    res["_fallbacks"] = dict(ops.fallback_counts)
    return res


def check_moe_layer():
    """MoE layer on the fused path (device-side layout + grouped GEMMs, bf16) vs the alltoall dispatcher with one module
    per expert (same weights, same kernels for the dense math): output, input gradient, router and expert weight grads."""
    import torch
    import torch.nn.functional as F
    from megatronapp_b200 import ops
    from megatronapp_b200.config.transformer_config import TransformerConfig
    from megatronapp_b200.models.gpt.gpt_layer_specs import get_mlp_module_spec
    from megatronapp_b200.parallel import state as ps
    from megatronapp_b200.parallel.random import model_parallel_cuda_manual_seed
    from megatronapp_b200.transformer.spec_utils import build_module

    ps.initialize_model_parallel(1, 1)
    model_parallel_cuda_manual_seed(123)
    H, FF, E = 512, 1024, 8
    outs = {}
    for kind, grouped in (("alltoall", False), ("fused", True)):
        cfg = TransformerConfig(num_layers=1, hidden_size=H, num_attention_heads=4, ffn_hidden_size=FF, num_moe_experts=E,
                                moe_router_topk=2, moe_token_dispatcher_type=kind, add_bias_linear=False,
                                activation_func=F.silu, gated_linear_unit=True, moe_grouped_gemm=grouped, bf16=True,
                                params_dtype=torch.bfloat16, moe_aux_loss_coeff=0.0)
        layer = build_module(get_mlp_module_spec(False, E, grouped), config=cfg).cuda()
        with torch.no_grad():
            torch.manual_seed(11)
            layer.router.weight.copy_(torch.randn_like(layer.router.weight.float()) * 0.3)
            for e in range(E):
                g = torch.Generator().manual_seed(1000 + e)
                w1 = (torch.randn(2 * FF, H, generator=g) * 0.05).bfloat16().cuda()
                w2 = (torch.randn(H, FF, generator=g) * 0.05).bfloat16().cuda()
                if grouped:
                    layer.experts.weight1[e * 2 * FF:(e + 1) * 2 * FF].copy_(w1)
                    layer.experts.weight2[e * H:(e + 1) * H].copy_(w2)
                else:
                    layer.experts.local_experts[e].linear_fc1.weight.copy_(w1)
                    layer.experts.local_experts[e].linear_fc2.weight.copy_(w2)
        g = torch.Generator().manual_seed(77)
        x = (torch.randn(1024, 2, H, generator=g) * 0.5).bfloat16().cuda().requires_grad_(True)
        out, _ = layer(x)
        (out.float() * torch.randn(out.shape, generator=torch.Generator().manual_seed(5)).cuda()).sum().backward()
        if grouped:
            w1g = layer.experts.weight1.grad.float()
            w2g = layer.experts.weight2.grad.float()
        else:
            w1g = torch.cat([ex.linear_fc1.weight.grad.float() for ex in layer.experts.local_experts])
            w2g = torch.cat([ex.linear_fc2.weight.grad.float() for ex in layer.experts.local_experts])
        outs[kind] = (out.detach(), x.grad.detach(), layer.router.weight.grad.detach(), w1g, w2g)
    a, b = outs["alltoall"], outs["fused"]
    res = {"out": _rel_err(b[0], a[0]), "dx": _rel_err(b[1], a[1]), "drouter": _rel_err(b[2], a[2]),
           "dw1": _rel_err(b[3], a[3]), "dw2": _rel_err(b[4], a[4]), "_fallbacks": dict(ops.fallback_counts)}
    return res


def check_grouped_perf():
    """Grouped expert GEMMs at Mixtral-8x7B shapes (h 4096, ffn 14336, SwiGLU, top-2) for a few expert-parallel splits."""
    import torch
    from megatronapp_b200 import ops
    from megatronapp_b200.ops import moe as mops

    res = {}
    dt = torch.bfloat16
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, E, rows_per_e, H, F1 in (("ep4_e2", 2, 8192, 4096, 28672), ("ep1_e8", 8, 2048, 4096, 28672),
                                       ("ep4_e2_etp2", 2, 8192, 4096, 14336)):
        counts = [rows_per_e + (37 * (i + 1)) % 200 - 100 for i in range(E)]
        lay = _ragged_layout(counts)
        rows = sum(counts)
        x = torch.randn(lay.cap, H, device="cuda", dtype=dt) * 0.5
        w1 = torch.randn(E * F1, H, device="cuda", dtype=dt) * 0.02
        dy = torch.randn(lay.cap, F1, device="cuda", dtype=dt) * 0.5
        dw = torch.zeros(E * F1, H, device="cuda", dtype=torch.float32)
        fl = 2.0 * rows * H * F1
        t = _time(lambda: mops.grouped_gemm(x, w1, lay, F1), flush=flush)
        res[f"{name}_fc1_fwd_tflops"] = fl / t / 1e9
        t = _time(lambda: mops.grouped_gemm(dy, w1, lay, F1, b_mn=True), flush=flush)
        res[f"{name}_fc1_dgrad_tflops"] = fl / t / 1e9
        t = _time(lambda: ops.ext().gemm_grouped(dy, x, dw, lay.offsets, 2, True, None, 0, None, None, E, F1), flush=flush)
        res[f"{name}_fc1_wgrad_tflops"] = fl / t / 1e9
        # per-expert loop of plain GEMMs on the same data (round-1 path) and cuBLAS for reference
        offs = lay.offsets.tolist()

        def loop():
            for g in range(E):
                ops.gemm(x[offs[g]:offs[g] + counts[g]], w1[g * F1:(g + 1) * F1])

        res[f"{name}_fc1_fwd_loop_tflops"] = fl / _time(loop, flush=flush) / 1e9

        def cublas():
            for g in range(E):
                torch.matmul(x[offs[g]:offs[g] + counts[g]], w1[g * F1:(g + 1) * F1].t())

        res[f"{name}_fc1_fwd_cublas_loop_tflops"] = fl / _time(cublas, flush=flush) / 1e9
        del x, w1, dy, dw
    return res


def check_smoke():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge

    ge.smoke()
    return {"ok": True}


def check_model_parity():
    """bf16 GPT step on the native kernels vs the fp32 PyTorch reference path (same weights)."""
    import torch
    from megatronapp_b200 import ops
    from megatronapp_b200.bench_model import build_gpt_for_bench, synthetic_batch, train_step
    from megatronapp_b200.parallel import state as ps

    st = build_gpt_for_bench(2, 512, 8, 256, 4096, 2, ffn=2048, device="cuda")
    b = synthetic_batch(st, "cuda")
    sd = {k: v.detach().float().cpu() for k, v in st.model[0].module.module.state_dict().items()}
    l_native = [train_step(st, [b, b]) for _ in range(3)]
    ps.destroy_model_parallel()
    st2 = build_gpt_for_bench(2, 512, 8, 256, 4096, 2, ffn=2048, device="cpu", dtype=torch.float32)
    st2.model[0].module.load_state_dict(sd)
    st2.optimizer.reload_model_params()
    b2 = {k: v.cpu() for k, v in b.items()}
    l_ref = [train_step(st2, [b2, b2]) for _ in range(3)]
    return {"native": l_native, "ref_fp32_cpu": l_ref, "fallbacks": dict(ops.fallback_counts),
            "max_abs_diff": max(abs(a - c) for a, c in zip(l_native, l_ref))}


def _paged_case(lens_new, lens_before, np_, ng, d, bs, dtype, seed=0):
    """A ragged step: request i already holds ``lens_before[i]`` keys and feeds ``lens_new[i]`` tokens.  Returns the
    kernel inputs and the fp32 reference output."""
    import torch

    torch.manual_seed(seed)
    dev = "cuda"
    R = len(lens_new)
    total = [b + n for b, n in zip(lens_before, lens_new)]
    nblk = [-(-t // bs) for t in total]
    n_blocks = sum(nblk) + 3
    perm = torch.randperm(n_blocks)[: sum(nblk)].tolist()              # pages scattered over the pool
    tables, o = [], 0
    for nb in nblk:
        tables.append(perm[o:o + nb])
        o += nb
    table = torch.zeros(R, max(nblk), dtype=torch.int32)
    for i, t in enumerate(tables):
        table[i, :len(t)] = torch.tensor(t, dtype=torch.int32)
    kc = torch.zeros(n_blocks * bs, ng, d, device=dev, dtype=dtype)
    vc = torch.zeros_like(kc)
    ks = [torch.randn(t, ng, d, device=dev, dtype=dtype) for t in total]
    vs = [torch.randn(t, ng, d, device=dev, dtype=dtype) for t in total]
    for i in range(R):
        pos = torch.arange(total[i])
        slot = torch.tensor(tables[i])[pos // bs] * bs + pos % bs
        kc[slot.to(dev)] = ks[i]
        vc[slot.to(dev)] = vs[i]
    T = sum(lens_new)
    q = torch.randn(T, np_, d, device=dev, dtype=dtype)
    tok_req, tok_len, ref = [], [], []
    scale = d ** -0.5
    qo = 0
    for i in range(R):
        for j in range(lens_new[i]):
            L = lens_before[i] + j + 1
            tok_req.append(i)
            tok_len.append(L)
            kk = ks[i][:L].float().repeat_interleave(np_ // ng, 1)       # [L, np, d]
            vv = vs[i][:L].float().repeat_interleave(np_ // ng, 1)
            sc = torch.einsum("hd,lhd->hl", q[qo].float(), kk) * scale
            ref.append(torch.einsum("hl,lhd->hd", torch.softmax(sc, -1), vv))
            qo += 1
    return (q, kc, vc, table.to(dev), torch.tensor(tok_req, dtype=torch.int32, device=dev),
            torch.tensor(tok_len, dtype=torch.int32, device=dev), bs, scale), torch.stack(ref)


def check_paged_attention():
    """paged-KV attention kernel vs an fp32 reference: decode + chunked-prefill mix, GQA groups 1/4/8, d 64/128, fp16/bf16,
    the flash-decoding split path, and the dynamic engine producing the same tokens with and without the kernel."""
    import torch

    from megatronapp_b200 import ops

    ext = ops.ext()
    out = {}
    for name, (new, before, np_, ng, d, bs, dt) in {
        "decode_g4_d128": ([1, 1, 1, 1, 1], [37, 512, 1, 200, 1023], 8, 2, 128, 16, torch.bfloat16),
        "mixed_g8_d128": ([1, 7, 1, 33], [300, 0, 45, 64], 16, 2, 128, 16, torch.bfloat16),
        "mha_d64_fp16": ([1, 5, 1], [129, 3, 77], 4, 4, 64, 32, torch.float16),
        "g2_d64": ([2, 1], [50, 260], 4, 2, 64, 16, torch.bfloat16),
    }.items():
        args, ref = _paged_case(new, before, np_, ng, d, bs, dt)
        for splits in (1, 4):
            got = ext.paged_attention(*args, splits)
            out[f"paged_{name}_s{splits}"] = _rel_err(got, ref)
    # engine parity: same greedy continuation through the kernel and through the gather + einsum path
    import os

    from megatronapp_b200.inference.dynamic import DynamicInferenceEngine
    from megatronapp_b200.inference.core import SamplingParams
    from megatronapp_b200.bench_model import build_gpt_for_bench

    from megatronapp_b200.data.tokenizer import NullTokenizer

    st = build_gpt_for_bench(2, 256, 4, 256, 512, 1, ffn=512, tp=1, device="cuda", kv_groups=2, rope=True, bias=False)
    model = st.model[0]
    model.eval()
    res = {}
    prompts = ["5 9 23 7 100 3 4 8 15 16 23 42", "11 3", " ".join(str(7 + 3 * i) for i in range(40))]
    for flag in ("1", "0"):
        os.environ["B200_PAGED_ATTN"] = flag
        eng = DynamicInferenceEngine(model, NullTokenizer(511), max_sequence_length=256, num_blocks=64, block_size=16,
                                     max_tokens_per_step=32, max_requests=3)
        eng.padded_vocab_size = 512
        with torch.no_grad():
            rs = eng.generate(prompts, SamplingParams(num_tokens_to_generate=12, top_k=1))
        res[flag] = ([[int(x) for x in r.generated_tokens] for r in rs], eng.context.kernel_calls)
    os.environ["B200_PAGED_ATTN"] = "1"
    out["paged_engine_token_mismatch"] = float(res["1"][0] != res["0"][0])
    out["paged_engine_kernel_unused"] = float(res["1"][1] == 0 or res["0"][1] != 0)
    return out


def check_paged_attention_perf():
    """decode step: 64 requests x 4096 keys, 32 q heads / 8 kv heads, d 128 -> achieved HBM bandwidth over the K/V bytes"""
    import torch

    from megatronapp_b200 import ops

    ext = ops.ext()
    out = {}
    for name, R, L, np_, ng in (("r64_l4096", 64, 4096, 32, 8), ("r8_l16384", 8, 16384, 32, 8), ("r256_l1024", 256, 1024, 32, 8)):
        args, _ = None, None
        torch.manual_seed(0)
        bs, d = 16, 128
        nblk = -(-L // bs)
        table = torch.randperm(R * nblk, dtype=torch.int32).view(R, nblk).cuda()
        kc = torch.randn(R * nblk * bs, ng, d, device="cuda", dtype=torch.bfloat16)
        vc = torch.randn_like(kc)
        q = torch.randn(R, np_, d, device="cuda", dtype=torch.bfloat16)
        tok_req = torch.arange(R, dtype=torch.int32, device="cuda")
        tok_len = torch.full((R,), L, dtype=torch.int32, device="cuda")
        ctas = R * ng
        best = None
        for splits in (1, 2, 4, 8, 16):
            ms = _time(lambda: ext.paged_attention(q, kc, vc, table, tok_req, tok_len, bs, d ** -0.5, splits))
            gbs = 2 * R * L * ng * d * 2 / ms / 1e6
            out[f"paged_{name}_s{splits}"] = {"ms": round(ms, 4), "GBs": round(gbs, 1)}
            best = max(best or 0, gbs)
        out[f"paged_{name}_best_frac_of_hbm"] = round(best / 7700.0, 3)

        def gather_ref():
            slots = (table.long().unsqueeze(-1) * bs + torch.arange(bs, device="cuda")).view(R, -1)[:, :L]
            kp, vp = kc[slots].repeat_interleave(np_ // ng, 2), vc[slots].repeat_interleave(np_ // ng, 2)
            s_ = torch.einsum("rhd,rlhd->rhl", q.float(), kp.float()) * d ** -0.5
            return torch.einsum("rhl,rlhd->rhd", torch.softmax(s_, -1), vp.float())

        if R * L <= 64 * 4096:
            out[f"paged_{name}_gather_einsum_ms"] = round(_time(gather_ref, iters=5, warmup=2), 3)
    return out


def check_long_context_causality():
    """Long sequences + grouped-query attention, the Llama-3 8B TP8 shape (4 query heads on 1 KV head, d 128, s 8192):
    (a) numerics of the native attention forward / backward vs the fp32 reference at s = 8192;
    (b) causality of a whole Llama-style model: changing the tokens after position p must leave every logit before p
        bit-identical (any leak -- attention mask, RoPE, sequence-parallel plumbing -- shows up here)."""
    import torch

    from megatronapp_b200 import ops
    from megatronapp_b200.bench_model import build_gpt_for_bench

    res = {}
    b, h, hk, s, d = 1, 4, 1, 8192, 128
    q, k, v = (t.detach().requires_grad_(True) for t in _attn_inputs(b, h, hk, s, d, True))
    o = ops._AttnFn.apply(q, k, v, 1.0 / d ** 0.5, True)
    do = torch.randn_like(o)
    o.backward(do)
    torch.cuda.synchronize()
    qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    orf = ops.attention_ref(qf, kf, vf, None, True)
    orf.backward(do.float())
    tag = f"long_attn_b{b}h{h}kv{hk}s{s}d{d}"
    res[tag + "_fwd"] = _rel_err(o, orf)
    res[tag + "_dq"], res[tag + "_dk"], res[tag + "_dv"] = _rel_err(q.grad, qf.grad), _rel_err(k.grad, kf.grad), _rel_err(v.grad, vf.grad)
    # first half of the outputs must not depend on the second half of the keys / values
    k2, v2 = k.detach().clone(), v.detach().clone()
    k2[:, :, s // 2:] = torch.randn_like(k2[:, :, s // 2:])
    v2[:, :, s // 2:] = torch.randn_like(v2[:, :, s // 2:])
    o2, _, _ = ops.ext().attn_fwd(q.detach(), k2, v2, 1.0 / d ** 0.5, True, False)
    res["long_attn_future_keys_leak"] = float((o2[:, :, : s // 2].float() - o.detach()[:, :, : s // 2].float()).abs().max())
    del q, k, v, qf, kf, vf, orf, o, o2
    torch.cuda.empty_cache()
    S = 4096
    st = build_gpt_for_bench(2, 512, 4, S, 4096, 1, ffn=1024, tp=1, device="cuda", swiglu=True, rope=True, rmsnorm=True, kv_groups=1,
                             bias=False)
    model = st.model[0]
    model.eval()
    g = torch.Generator().manual_seed(5)
    toks = torch.randint(1, 4096, (1, S), generator=g).cuda()
    pos = torch.arange(S, device="cuda").unsqueeze(0)
    with torch.no_grad():
        a = model(toks, pos, None).float()
        t2 = toks.clone()
        cut = S // 2 + 37
        t2[:, cut:] = torch.randint(1, 4096, (1, S - cut), generator=g).cuda()
        bb = model(t2, pos, None).float()
    if a.shape[0] == S:                                   # [s, b, v] -> [b, s, v]
        a, bb = a.transpose(0, 1), bb.transpose(0, 1)
    res["model_future_tokens_leak"] = float((a[:, :cut] - bb[:, :cut]).abs().max())
    res["model_suffix_changed"] = float((a[:, cut:] - bb[:, cut:]).abs().max() > 0)
    return res


CHECKS = {
    "gemm_nt": lambda: check_gemm("nt"),
    "gemm_nt_bias_gelu": lambda: check_gemm("nt_bias_gelu", [(256, 256, 128), (2048, 4096, 4096), (300, 520, 264)]),
    "gemm_nn": lambda: check_gemm("nn"),
    "gemm_tn": lambda: check_gemm("tn"),
    "gemm_tt": lambda: check_gemm("tt"),
    "gemm_tt_accum": lambda: check_gemm("tt_accum"),
    "gemm_nt_fp16": lambda: check_gemm("nt", [(256, 256, 128), (2048, 4096, 4096)], "fp16"),
    "elementwise": check_elementwise,
    "attention_fwd": check_attention_fwd,
    "attention": check_attention,
    "fp8": check_fp8,
    "mxfp8": check_mxfp8,
    "grouped_gemm": check_grouped_gemm,
    "moe_kernels": check_moe_kernels,
    "moe_layer": check_moe_layer,
    "paged_attention": check_paged_attention,
    "long_context_causality": check_long_context_causality,
    "model_parity": check_model_parity,
    "smoke": check_smoke,
}
PERF = {"paged_attention_perf": check_paged_attention_perf, "mxfp8_perf": check_mxfp8_perf, "grouped_perf": check_grouped_perf, "fp8_perf": check_fp8_perf, "gemm_perf": check_gemm_perf, "elementwise_perf": check_elementwise_perf, "attention_perf": check_attention_perf}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "checks.json"))
    ap.add_argument("--only", default=None)
    ap.add_argument("--perf", action="store_true")
    ap.add_argument("--one", default=None)
    ap.add_argument("--timeout", type=int, default=240)
    a = ap.parse_args()
    allc = dict(CHECKS)
    allc.update(PERF)
    if a.one:
        r = allc[a.one]()
        print("RESULT_JSON " + json.dumps(r))
        return
    names = a.only.split(",") if a.only else list(CHECKS) + (list(PERF) if a.perf else [])
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    results = {}
    if os.path.exists(a.out):
        try:
            results = json.load(open(a.out))
        except Exception:
            results = {}
    for n in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True,
                               text=True, timeout=a.timeout, cwd=ROOT)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT_JSON ")]
            if line:
                results[n] = json.loads(line[-1][len("RESULT_JSON "):])
            else:
                results[n] = {"error": f"rc={p.returncode}", "stderr": p.stderr[-1500:], "stdout": p.stdout[-500:]}
        except subprocess.TimeoutExpired:
            results[n] = {"error": f"TIMEOUT after {a.timeout}s"}
        results[n + "_secs"] = round(time.time() - t0, 1)
        json.dump(results, open(a.out, "w"), indent=1)
        print(n, json.dumps(results[n])[:1500], flush=True)


if __name__ == "__main__":
    main()
