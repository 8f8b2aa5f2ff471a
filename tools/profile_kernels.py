"""Tiny driver for ncu captures: one launch of each headline kernel at the 6.7B benchmark shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from megatronapp_b200 import ops
from megatronapp_b200.ops import fp8

dt = torch.bfloat16
x = torch.randn(2048, 4096, device="cuda", dtype=dt)
w1 = torch.randn(16384, 4096, device="cuda", dtype=dt)
bias = torch.randn(16384, device="cuda", dtype=dt)
g = torch.ones(4096, device="cuda", dtype=dt)
b = torch.zeros(4096, device="cuda", dtype=dt)
dy = torch.randn(2048, 16384, device="cuda", dtype=dt)
main = torch.zeros(16384, 4096, device="cuda", dtype=torch.float32)
q, k, v = (torch.randn(2048, 4, 32, 128, device="cuda", dtype=dt).permute(1, 2, 0, 3) for _ in range(3))
x8, _, sx, _ = fp8.quantize(x)
w8, _, sw, _ = fp8.quantize(w1)
xm, xmc = fp8.quantize_mx_native(x, False, 0, 1)
wm, _ = fp8.quantize_mx_native(w1, False, 1, -1)
# MoE: 8 experts x ~2048 rows of the ragged 256-row padded layout (Mixtral expert width)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from gpu_checks import _ragged_layout  # noqa: E402
from megatronapp_b200.ops import moe as mops  # noqa: E402

lay = _ragged_layout([2048 + 37 * i for i in range(8)])
xe = torch.randn(lay.cap, 4096, device="cuda", dtype=dt)
we = torch.randn(8 * 7168, 4096, device="cuda", dtype=dt) * 0.02
dye = torch.randn(lay.cap, 7168, device="cuda", dtype=dt)
dwe = torch.zeros(8 * 7168, 4096, device="cuda", dtype=torch.float32)
for _ in range(3):
    pre = torch.empty(2048, 16384, device="cuda", dtype=dt)
    ops.gemm(x, w1, bias=bias, act=1, out_pre=pre)                       # fc1 fwd, 2-CTA kernel + bias/GELU epilogue
    ops.gemm(dy, w1, b_mn=True)                                          # fc1 dgrad
    ops.gemm(dy, x, out=main, a_mn=True, b_mn=True, epilogue=1)          # fc1 wgrad, TMA reduce-add epilogue
    ops.gemm(x, w1, block_n=256)                                         # 1-CTA kernel
    fp8.gemm_fp8(x8, w8, sx, sw)                                         # fp8 kernel
    fp8.gemm_mx_native(xm, wm)                                           # block-scaled mxfp8 kernel (UTCCP + UTCQMMA)
    fp8.quantize_mx_native(x, False, 0, 1)                               # dual row/column MX quantiser
    mops.grouped_gemm(xe, we, lay, 7168)                                 # grouped expert GEMM (forward)
    mops.grouped_gemm(dye, we, lay, 7168, b_mn=True)                     # grouped dgrad
    ops.ext().gemm_grouped(dye, xe, dwe, lay.offsets, 2, True, None, 0, None, None, 8, 7168)   # grouped wgrad
    mops._gather_rows(x, lay, 8)                                         # permute
    mops._gather_sum(xe, lay)                                            # combine
    o, lse, _ = ops.ext().attn_fwd(q, k, v, 128 ** -0.5, True, False)
    ops.ext().attn_bwd(torch.randn_like(o), q, k, v, o, lse, 128 ** -0.5, True, None, None, None)
    y, mean, rstd = ops.ext().layernorm_fwd(x, g, b, 1e-5, False, False)
    ops.ext().layernorm_bwd(x, x, g, mean, rstd, False, False, True)
torch.cuda.synchronize()
