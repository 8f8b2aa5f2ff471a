"""Tiny driver for ncu captures of the attention kernels (fwd + both bwd kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from megatronapp_b200 import ops

b, h, s, d = int(os.environ.get("B", 2)), 32, 2048, 128
q, k, v = (torch.randn(s, b, h, d, device="cuda", dtype=torch.bfloat16).permute(1, 2, 0, 3) for _ in range(3))
for _ in range(3):
    o, lse, _ = ops.ext().attn_fwd(q, k, v, d ** -0.5, True, False)
    ops.ext().attn_bwd(torch.randn_like(o), q, k, v, o, lse, d ** -0.5, True, None, None, None)
torch.cuda.synchronize()
