"""RADIO ViT encoder (reference ``core/models/vision/radio.py:22``): a ViT whose position table is laid out on the
*largest* supported patch grid and resampled to whatever resolution arrives, with several summary ("class") tokens.

Differences from ``CLIPViTModel``: patches go through a linear embedder (column-parallel) instead of a conv, position
embeddings cover patches only, and come in three flavours -- exact (input grid == max grid), windowed / bilinearly
resized at eval, and *cropped position embeddings* (CPE) during training, where every sample looks at a random
scale / aspect / offset window of the table so the model learns resolution-agnostic positions.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...parallel.tp_layers import ColumnParallelLinear
from ...transformer.module import MegatronModule
from ...transformer.spec_utils import ModuleSpec
from ...transformer.transformer_block import TransformerBlock
from .clip_vit_model import get_vit_layer_spec


def patchify(x: torch.Tensor, p: int) -> torch.Tensor:
    """``[b, c, H, W] -> [b, (H/p)(W/p), c*p*p]`` with (c, y, x) minor order."""
    b, c, H, W = x.shape
    return x.view(b, c, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(b, (H // p) * (W // p), c * p * p)


class RADIOViTModel(MegatronModule):
    def __init__(self, transformer_config, transformer_layer_spec: Optional[ModuleSpec] = None, ln_pre_impl=None,
                 ln_post_impl=None, use_mask_token: bool = False, add_class_token: bool = True, class_token_len: int = 8,
                 patch_dim: int = 16, img_h: int = 224, img_w: int = 224, max_img_h: int = 2048, max_img_w: int = 2048,
                 pos_dropout: float = 0.0, has_cpe: bool = True, embedder_bias: bool = False):
        super().__init__(transformer_config)
        cfg = transformer_config
        self.visual_hidden_size, self.patch_dim = cfg.hidden_size, patch_dim
        self.img_h, self.img_w = img_h, img_w
        self.input_dims = (img_h // patch_dim, img_w // patch_dim)
        self.max_rows, self.max_cols = max_img_h // patch_dim, max_img_w // patch_dim
        self.add_class_token, self.class_token_len = add_class_token, class_token_len
        self.pos_dropout, self.has_cpe = pos_dropout, has_cpe
        dt = cfg.params_dtype
        if use_mask_token:
            self.mask_token = nn.Parameter(torch.zeros(1, cfg.hidden_size, dtype=dt))
        if add_class_token:
            self.class_token = nn.Parameter(torch.randn(class_token_len, cfg.hidden_size, dtype=dt))
        self.seq_length = self.input_dims[0] * self.input_dims[1] + (class_token_len if add_class_token else 0)
        self.position_embeddings = nn.Parameter(torch.randn(1, self.max_rows * self.max_cols, cfg.hidden_size, dtype=dt) * 0.02)
        self.embedder = ColumnParallelLinear(3 * patch_dim * patch_dim, cfg.hidden_size, config=cfg, init_method=cfg.init_method,
                                             gather_output=True, bias=embedder_bias, skip_bias_add=False)
        self.ln_pre = ln_pre_impl(config=cfg, hidden_size=cfg.hidden_size, eps=cfg.layernorm_epsilon) if ln_pre_impl else None
        self.ln_post = ln_post_impl(config=cfg, hidden_size=cfg.hidden_size, eps=cfg.layernorm_epsilon) if ln_post_impl else None
        self.decoder = TransformerBlock(cfg, transformer_layer_spec or get_vit_layer_spec(), pre_process=True, post_process=False)

    def set_input_tensor(self, input_tensor):
        self.decoder.set_input_tensor(input_tensor)

    # -- position table -----------------------------------------------------------
    def _table(self) -> torch.Tensor:
        return self.position_embeddings.view(1, self.max_rows, self.max_cols, -1).permute(0, 3, 1, 2)    # [1, h, R, C]

    def _cpe_grid(self, b: int, dims: Tuple[int, int], dev) -> torch.Tensor:
        """Sampling grid of a random window: area scale in [0.1, 1], aspect in [3/4, 4/3], uniform offset."""
        lo = math.sqrt(0.1)
        scale = torch.rand(b, 1, 1, device=dev) * (1 - lo) + lo
        amax = math.log(4 / 3)
        aspect = torch.exp((torch.rand(b, 1, 1, device=dev) * 2 - 1) * amax)
        sxy = torch.stack([scale * aspect, scale / aspect], dim=-1).clamp_(0, 1)                       # [b,1,1,2]
        off = torch.rand(b, 1, 1, 2, device=dev) * (1 - sxy)
        gx = torch.linspace(0, 1, dims[1], device=dev)[None, None, :].expand(b, dims[0], -1)
        gy = torch.linspace(0, 1, dims[0], device=dev)[None, :, None].expand(b, -1, dims[1])
        return (torch.stack([gx, gy], dim=-1) * sxy + off) * 2 - 1

    def get_pos_enc(self, batch_size: int, patch_idxs: Optional[torch.Tensor] = None,
                    input_size: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        dims = self.input_dims if input_size is None else tuple(d // self.patch_dim for d in input_size)
        if dims == (self.max_rows, self.max_cols):
            pe = self.position_embeddings
        else:
            t = self._table()
            if self.has_cpe and self.training:
                t = F.grid_sample(t.float().expand(batch_size, -1, -1, -1), self._cpe_grid(batch_size, dims, t.device),
                                  mode="bilinear", padding_mode="zeros", align_corners=True).to(t.dtype)
            else:
                if self.has_cpe:       # eval: shrink the table to a square of the longer side, then crop the shorter one
                    m = max(dims)
                    t = F.interpolate(t.float(), size=(m, m), mode="bilinear", align_corners=True).to(t.dtype)
                t = t[..., :min(dims[0], t.shape[-2]), :min(dims[1], t.shape[-1])]
            if tuple(t.shape[-2:]) != dims:
                t = F.interpolate(t.float(), size=dims, mode="bilinear", align_corners=True).to(t.dtype)
            pe = t.flatten(2).permute(0, 2, 1)
        if patch_idxs is not None:
            pe = torch.gather(pe.expand(patch_idxs.shape[0], -1, -1), 1, patch_idxs.unsqueeze(-1).expand(-1, -1, pe.shape[-1]))
        return pe

    def apply_pos_enc(self, patches, patch_idxs=None, input_size=None):
        pe = self.get_pos_enc(patches.shape[0], patch_idxs, input_size)
        if self.training and self.pos_dropout > 0:
            keep = torch.rand(patches.shape[0], 1, 1, device=pe.device) > self.pos_dropout
            return patches + torch.where(keep, pe, torch.zeros_like(pe)), pe
        return patches + pe, pe

    def forward(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``x [b, 3, H, W]`` (any H, W that are multiples of the patch size) -> ``[b, cls + patches, hidden]``."""
        size = x.shape[2:]
        x, _ = self.embedder(patchify(x.to(self.embedder.weight.dtype), self.patch_dim))
        x, _ = self.apply_pos_enc(x, input_size=size)
        if self.add_class_token:
            x = torch.cat([self.class_token.unsqueeze(0).expand(x.shape[0], -1, -1).to(x.dtype), x], dim=1)
        if self.ln_pre is not None:
            x = self.ln_pre(x)
        x = self.decoder(x.permute(1, 0, 2).contiguous(), attention_mask)
        x = x.permute(1, 0, 2).contiguous()
        return self.ln_post(x) if self.ln_post is not None else x
