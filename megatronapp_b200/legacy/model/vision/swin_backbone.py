"""Swin Transformer backbone (reference ``legacy/model/vision/swin_backbone.py``, 625 lines): hierarchical stages of
(shifted-)window multi-head attention with a learned relative position bias, patch merging between stages."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


def window_partition(x: torch.Tensor, ws: int) -> torch.Tensor:
    """[B, H, W, C] -> [B * nW, ws, ws, C]"""
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)


def window_reverse(w: torch.Tensor, ws: int, H: int, W: int) -> torch.Tensor:
    B = w.shape[0] // ((H // ws) * (W // ws))
    x = w.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


class WindowAttention(nn.Module):
    def __init__(self, dim: int, ws: int, heads: int, qkv_bias: bool = True, attn_drop: float = 0.0, proj_drop: float = 0.0):
        super().__init__()
        self.dim, self.ws, self.heads = dim, ws, heads
        self.scale = (dim // heads) ** -0.5
        self.rel_bias = nn.Parameter(torch.zeros((2 * ws - 1) ** 2, heads))
        nn.init.trunc_normal_(self.rel_bias, std=0.02)
        c = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
        rel = (c[:, :, None] - c[:, None, :]).permute(1, 2, 0) + (ws - 1)
        self.register_buffer("rel_index", rel[..., 0] * (2 * ws - 1) + rel[..., 1], persistent=False)
        self.qkv = nn.Linear(dim, 3 * dim, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.attn_drop, self.proj_drop = attn_drop, nn.Dropout(proj_drop)

    def forward(self, x: torch.Tensor, mask: torch.Tensor = None) -> torch.Tensor:
        Bn, N, C = x.shape
        q, k, v = self.qkv(x).view(Bn, N, 3, self.heads, C // self.heads).permute(2, 0, 3, 1, 4)
        bias = self.rel_bias[self.rel_index.view(-1)].view(N, N, -1).permute(2, 0, 1).unsqueeze(0)
        if mask is not None:                          # [nW, N, N] additive mask of the shifted windows
            nW = mask.shape[0]
            bias = bias + mask.view(1, nW, 1, N, N).expand(Bn // nW, -1, -1, -1, -1).reshape(Bn, 1, N, N)
        out = F.scaled_dot_product_attention(q, k, v, attn_mask=bias.to(q.dtype), dropout_p=self.attn_drop if self.training else 0.0,
                                             scale=self.scale)
        return self.proj_drop(self.proj(out.transpose(1, 2).reshape(Bn, N, C)))


class SwinBlock(nn.Module):
    def __init__(self, dim, res, heads, ws=7, shift=0, mlp_ratio=4.0, drop=0.0, attn_drop=0.0, drop_path=0.0):
        super().__init__()
        self.res, self.ws, self.shift = res, min(ws, min(res)), shift if min(res) > ws else 0
        self.norm1, self.norm2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, self.ws, heads, attn_drop=attn_drop, proj_drop=drop)
        self.mlp = nn.Sequential(nn.Linear(dim, int(dim * mlp_ratio)), nn.GELU(), nn.Dropout(drop),
                                 nn.Linear(int(dim * mlp_ratio), dim), nn.Dropout(drop))
        self.drop_path = drop_path
        mask = None
        if self.shift > 0:
            H, W = res
            img = torch.zeros(1, H, W, 1)
            cnt = 0
            for hs in (slice(0, -self.ws), slice(-self.ws, -self.shift), slice(-self.shift, None)):
                for wsl in (slice(0, -self.ws), slice(-self.ws, -self.shift), slice(-self.shift, None)):
                    img[:, hs, wsl, :] = cnt
                    cnt += 1
            mw = window_partition(img, self.ws).view(-1, self.ws * self.ws)
            mask = (mw.unsqueeze(1) - mw.unsqueeze(2)).ne(0).float() * -100.0
        self.register_buffer("attn_mask", mask, persistent=False)

    def _dp(self, x):
        if not self.training or self.drop_path == 0.0:
            return x
        keep = torch.rand(x.shape[0], 1, 1, device=x.device) >= self.drop_path
        return x * keep / (1.0 - self.drop_path)

    def forward(self, x):
        H, W = self.res
        B, L, C = x.shape
        y = self.norm1(x).view(B, H, W, C)
        if self.shift:
            y = torch.roll(y, shifts=(-self.shift, -self.shift), dims=(1, 2))
        y = self.attn(window_partition(y, self.ws).view(-1, self.ws * self.ws, C), self.attn_mask)
        y = window_reverse(y.view(-1, self.ws, self.ws, C), self.ws, H, W)
        if self.shift:
            y = torch.roll(y, shifts=(self.shift, self.shift), dims=(1, 2))
        x = x + self._dp(y.reshape(B, L, C))
        return x + self._dp(self.mlp(self.norm2(x)))


class PatchMerging(nn.Module):
    def __init__(self, res, dim):
        super().__init__()
        self.res, self.norm, self.reduction = res, nn.LayerNorm(4 * dim), nn.Linear(4 * dim, 2 * dim, bias=False)

    def forward(self, x):
        H, W = self.res
        B, L, C = x.shape
        x = x.view(B, H, W, C)
        x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], dim=-1)
        return self.reduction(self.norm(x.view(B, -1, 4 * C)))


class SwinTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, depths: Sequence[int] = (2, 2, 6, 2),
                 num_heads: Sequence[int] = (3, 6, 12, 24), window_size=7, mlp_ratio=4.0, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.1, output_avg=False):
        super().__init__()
        self.embed = nn.Conv2d(in_chans, embed_dim, patch_size, patch_size)
        self.embed_norm = nn.LayerNorm(embed_dim)
        res = [img_size // patch_size, img_size // patch_size]
        dpr = torch.linspace(0, drop_path_rate, sum(depths)).tolist()
        self.stages, self.merges = nn.ModuleList(), nn.ModuleList()
        dim, k = embed_dim, 0
        self.out_dims: List[int] = []
        for i, (d, nh) in enumerate(zip(depths, num_heads)):
            self.stages.append(nn.Sequential(*[SwinBlock(dim, tuple(res), nh, window_size, 0 if j % 2 == 0 else window_size // 2,
                                                         mlp_ratio, drop_rate, attn_drop_rate, dpr[k + j]) for j in range(d)]))
            k += d
            self.out_dims.append(dim)
            if i < len(depths) - 1:
                self.merges.append(PatchMerging(tuple(res), dim))
                dim, res = dim * 2, [res[0] // 2, res[1] // 2]
        self.norm = nn.LayerNorm(dim)
        self.num_features, self.output_avg = dim, output_avg

    def forward(self, x: torch.Tensor):
        """returns the per-stage feature maps ([B, C_i, H_i, W_i]) or, with ``output_avg``, the pooled final feature."""
        x = self.embed_norm(self.embed(x).flatten(2).transpose(1, 2))
        feats = []
        for i, st in enumerate(self.stages):
            x = st(x)
            H, W = st[0].res
            feats.append((self.norm(x) if i == len(self.stages) - 1 else x).transpose(1, 2).reshape(x.shape[0], -1, H, W))
            if i < len(self.merges):
                x = self.merges[i](x)
        return feats[-1].flatten(2).mean(-1) if self.output_avg else feats


_SWIN_TYPES = {   # reference esvit_swin_backbone.get_swin (--swin-backbone-type): embed dim, depths, heads, student drop-path
    "tiny": (96, (2, 2, 6, 2), (3, 6, 12, 24), 0.1),
    "base": (128, (2, 2, 18, 2), (4, 8, 16, 32), 0.2),
    "h3": (384, (2, 2, 18, 2), (6, 12, 24, 48), 0.2),
}


def get_swin(img_size=224, drop_path_rate=None, output_avg=False, tiny=True, kind=None, is_teacher=False, in_chans=3):
    """``kind`` in tiny | base | h3 (``tiny=`` kept for older callers); the teacher of a self-distillation pair runs without
    stochastic depth."""
    dim, depths, heads, dpr = _SWIN_TYPES[kind or ("tiny" if tiny else "base")]
    if drop_path_rate is None:
        drop_path_rate = dpr
    return SwinTransformer(img_size=img_size, window_size=7, drop_path_rate=0.0 if is_teacher else drop_path_rate,
                           output_avg=output_avg, embed_dim=dim, depths=depths, num_heads=heads, in_chans=in_chans)
