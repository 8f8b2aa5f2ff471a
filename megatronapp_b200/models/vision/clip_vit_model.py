"""CLIP / SigLIP style vision transformer encoder.

Parity: reference ``models/vision/clip_vit_model.py`` (``CLIPViTModel``: conv patch embedding,
optional class token, learned position embeddings, pre-LayerNorm, a ``TransformerBlock`` with
bidirectional attention, ``get_num_image_embeddings``) and ``vit_layer_specs.py``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ...transformer.attention import DotProductAttention, SelfAttention, SelfAttentionSubmodules
from ...transformer.mlp import MLP, MLPSubmodules
from ...transformer.module import MegatronModule
from ...transformer.norms import FusedLayerNorm
from ...transformer.spec_utils import ModuleSpec
from ...transformer.transformer_block import TransformerBlock
from ...transformer.transformer_layer import TransformerLayer, TransformerLayerSubmodules
from ...parallel.tp_layers import ColumnParallelLinear, RowParallelLinear


def get_vit_layer_spec() -> ModuleSpec:
    attn = ModuleSpec(module=SelfAttention, params={"attn_mask_type": "no_mask"},
                      submodules=SelfAttentionSubmodules(linear_qkv=ColumnParallelLinear, core_attention=DotProductAttention,
                                                         linear_proj=RowParallelLinear))
    mlp = ModuleSpec(module=MLP, submodules=MLPSubmodules(linear_fc1=ColumnParallelLinear, linear_fc2=RowParallelLinear))
    return ModuleSpec(module=TransformerLayer, submodules=TransformerLayerSubmodules(
        input_layernorm=FusedLayerNorm, self_attention=attn, pre_mlp_layernorm=FusedLayerNorm, mlp=mlp))


def get_num_image_embeddings(img_h: int, img_w: int, patch_dim: int, vision_model_type: str = "clip",
                             disable_vision_class_token: bool = False, class_token_len: int = 1,
                             pixel_shuffle: bool = False) -> int:
    n = (img_h // patch_dim) * (img_w // patch_dim)
    if pixel_shuffle:
        n //= 4
    if vision_model_type in ("clip", "radio") and not disable_vision_class_token:
        n += class_token_len
    return n


class CLIPViTModel(MegatronModule):
    def __init__(self, transformer_config, transformer_layer_spec: Optional[ModuleSpec] = None, ln_pre_impl=FusedLayerNorm,
                 ln_post_impl=None, add_class_token: bool = True, class_token_len: int = 1, patch_dim: int = 14,
                 img_h: int = 336, img_w: int = 336, model_subtype: str = "clip"):
        super().__init__(transformer_config)
        cfg = transformer_config
        self.visual_hidden_size = cfg.hidden_size
        self.patch_dim, self.img_h, self.img_w = patch_dim, img_h, img_w
        assert img_h % patch_dim == 0 and img_w % patch_dim == 0
        self.num_patches = (img_h // patch_dim) * (img_w // patch_dim)
        self.add_class_token, self.class_token_len = add_class_token, class_token_len
        self.seq_length = self.num_patches + (class_token_len if add_class_token else 0)
        dt = cfg.params_dtype
        self.conv1 = nn.Conv2d(3, cfg.hidden_size, kernel_size=patch_dim, stride=patch_dim,
                               bias=model_subtype == "siglip", dtype=dt)
        self.position_embeddings = nn.Embedding(self.seq_length, cfg.hidden_size, dtype=dt)
        self.register_buffer("position_ids", torch.arange(self.seq_length).unsqueeze(0), persistent=False)
        if add_class_token:
            self.class_token = nn.Parameter(torch.randn(1, class_token_len, cfg.hidden_size, dtype=dt) * 0.02)
        self.ln_pre = ln_pre_impl(config=cfg, hidden_size=cfg.hidden_size, eps=cfg.layernorm_epsilon) \
            if (ln_pre_impl is not None and model_subtype == "clip") else None
        self.decoder = TransformerBlock(cfg, transformer_layer_spec or get_vit_layer_spec(), pre_process=True,
                                        post_process=False)
        self.ln_post = ln_post_impl(config=cfg, hidden_size=cfg.hidden_size, eps=cfg.layernorm_epsilon) \
            if ln_post_impl is not None else None

    def set_input_tensor(self, input_tensor):
        self.decoder.set_input_tensor(input_tensor)

    def forward(self, x: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x: [b, 3, H, W] -> [b, seq, hidden]."""
        x = self.conv1(x.to(self.conv1.weight.dtype))                 # [b, h, gh, gw]
        x = x.flatten(2).permute(0, 2, 1)                              # [b, patches, h]
        if self.add_class_token:
            x = torch.cat([self.class_token.expand(x.shape[0], -1, -1).to(x.dtype), x], dim=1)
        x = x + self.position_embeddings(self.position_ids.to(x.device))
        if self.ln_pre is not None:
            x = self.ln_pre(x)
        x = x.permute(1, 0, 2).contiguous()                            # [s, b, h]
        x = self.decoder(x, attention_mask)
        x = x.permute(1, 0, 2).contiguous()
        if self.ln_post is not None:
            x = self.ln_post(x)
        return x
