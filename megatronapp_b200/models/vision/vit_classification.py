"""ViT image classification (reference ``megatron/legacy/model/vision/classification.py`` ``VitClassificationModel``
+ ``vit_backbone.py``): the CLIP-style ViT encoder with a class token and an MLP head on top of it."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...transformer.module import MegatronModule
from ...transformer.norms import FusedLayerNorm
from .clip_vit_model import CLIPViTModel


class VitClassificationModel(MegatronModule):
    def __init__(self, config, num_classes: int, img_h: int = 224, img_w: int = 224, patch_dim: int = 16,
                 finetune: bool = False, pre_process: bool = True, post_process: bool = True):
        super().__init__(config)
        self.num_classes, self.finetune = num_classes, finetune
        self.backbone = CLIPViTModel(config, None, ln_pre_impl=None, ln_post_impl=FusedLayerNorm, add_class_token=True,
                                     patch_dim=patch_dim, img_h=img_h, img_w=img_w)
        h = config.hidden_size
        if finetune:
            self.head = nn.Linear(h, num_classes, dtype=config.params_dtype)
            nn.init.zeros_(self.head.weight)
        else:
            self.head = nn.Sequential(nn.Linear(h, h, dtype=config.params_dtype), nn.Tanh(),
                                      nn.Linear(h, num_classes, dtype=config.params_dtype))
        self.share_embeddings_and_output_weights = False

    def set_input_tensor(self, t):
        self.backbone.set_input_tensor(t)

    def forward(self, images: torch.Tensor) -> torch.Tensor:
        hidden = self.backbone(images)          # [b, 1 + patches, h]
        return self.head(hidden[:, 0, :])
