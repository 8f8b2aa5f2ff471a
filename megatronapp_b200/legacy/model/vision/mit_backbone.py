"""Mix Transformer (SegFormer encoder; reference ``legacy/model/vision/mit_backbone.py``, 415 lines): overlapping patch
embeddings, efficient self-attention with spatial reduction of keys / values, Mix-FFN with a 3x3 depth-wise convolution."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


class OverlapPatchEmbed(nn.Module):
    def __init__(self, patch, stride, cin, dim):
        super().__init__()
        self.proj, self.norm = nn.Conv2d(cin, dim, patch, stride, patch // 2), nn.LayerNorm(dim)

    def forward(self, x):
        x = self.proj(x)
        H, W = x.shape[2:]
        return self.norm(x.flatten(2).transpose(1, 2)), H, W


class EfficientAttention(nn.Module):
    def __init__(self, dim, heads, sr_ratio, drop=0.0):
        super().__init__()
        self.heads, self.q, self.kv, self.proj = heads, nn.Linear(dim, dim), nn.Linear(dim, 2 * dim), nn.Linear(dim, dim)
        self.sr = nn.Conv2d(dim, dim, sr_ratio, sr_ratio) if sr_ratio > 1 else None
        self.norm = nn.LayerNorm(dim) if sr_ratio > 1 else None
        self.drop = drop

    def forward(self, x, H, W):
        B, N, C = x.shape
        q = self.q(x).view(B, N, self.heads, C // self.heads).transpose(1, 2)
        y = x
        if self.sr is not None:
            y = self.norm(self.sr(x.transpose(1, 2).reshape(B, C, H, W)).flatten(2).transpose(1, 2))
        k, v = self.kv(y).view(B, -1, 2, self.heads, C // self.heads).permute(2, 0, 3, 1, 4)
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=self.drop if self.training else 0.0)
        return self.proj(o.transpose(1, 2).reshape(B, N, C))


class MixFFN(nn.Module):
    def __init__(self, dim, hidden, drop=0.0):
        super().__init__()
        self.fc1, self.dw, self.fc2, self.drop = nn.Linear(dim, hidden), nn.Conv2d(hidden, hidden, 3, 1, 1, groups=hidden), nn.Linear(hidden, dim), nn.Dropout(drop)

    def forward(self, x, H, W):
        x = self.fc1(x)
        B, N, C = x.shape
        x = self.dw(x.transpose(1, 2).reshape(B, C, H, W)).flatten(2).transpose(1, 2)
        return self.drop(self.fc2(self.drop(F.gelu(x))))


class MitBlock(nn.Module):
    def __init__(self, dim, heads, mlp_ratio, sr_ratio, drop, drop_path):
        super().__init__()
        self.n1, self.n2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.attn, self.mlp, self.dp = EfficientAttention(dim, heads, sr_ratio, drop), MixFFN(dim, int(dim * mlp_ratio), drop), drop_path

    def _dp(self, x):
        if not self.training or self.dp == 0.0:
            return x
        return x * (torch.rand(x.shape[0], 1, 1, device=x.device) >= self.dp) / (1.0 - self.dp)

    def forward(self, x, H, W):
        x = x + self._dp(self.attn(self.n1(x), H, W))
        return x + self._dp(self.mlp(self.n2(x), H, W))


class MixVisionTransformer(nn.Module):
    def __init__(self, in_chans=3, embed_dims: Sequence[int] = (32, 64, 160, 256), num_heads=(1, 2, 5, 8), mlp_ratios=(4, 4, 4, 4),
                 depths=(2, 2, 2, 2), sr_ratios=(8, 4, 2, 1), drop_rate=0.0, drop_path_rate=0.1, output_avg=False):
        super().__init__()
        dpr = torch.linspace(0, drop_path_rate, sum(depths)).tolist()
        self.embeds, self.blocks, self.norms = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        k, cin = 0, in_chans
        for i, dim in enumerate(embed_dims):
            self.embeds.append(OverlapPatchEmbed(7 if i == 0 else 3, 4 if i == 0 else 2, cin, dim))
            self.blocks.append(nn.ModuleList([MitBlock(dim, num_heads[i], mlp_ratios[i], sr_ratios[i], drop_rate, dpr[k + j])
                                              for j in range(depths[i])]))
            self.norms.append(nn.LayerNorm(dim))
            k, cin = k + depths[i], dim
        self.embed_dims, self.output_avg = list(embed_dims), output_avg

    def forward(self, x):
        feats = []
        for emb, blks, norm in zip(self.embeds, self.blocks, self.norms):
            x, H, W = emb(x)
            for b in blks:
                x = b(x, H, W)
            x = norm(x).transpose(1, 2).reshape(x.shape[0], -1, H, W)
            feats.append(x)
        return feats[-1].flatten(2).mean(-1) if self.output_avg else feats


def mit_b0(**kw):
    return MixVisionTransformer(embed_dims=(32, 64, 160, 256), depths=(2, 2, 2, 2), **kw)


def mit_b1(**kw):
    return MixVisionTransformer(embed_dims=(64, 128, 320, 512), depths=(2, 2, 2, 2), **kw)


def mit_b2(**kw):
    return MixVisionTransformer(embed_dims=(64, 128, 320, 512), depths=(3, 4, 6, 3), **kw)
