"""Image inpainting (reference ``legacy/model/vision/inpainting.py`` + ``pretrain_vision_inpaint.py``): reconstruct the
masked region of an image.  ``VitInpaintingModel``: ViT trunk + per-patch linear decoder (un-patchify);
``MitInpaintingModel``: MiT encoder + an all-MLP (SegFormer-style) decoder that fuses the four scales."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class VitInpaintingModel(nn.Module):
    def __init__(self, config, img_h=224, img_w=224, patch_dim=16, pre_process=True, post_process=True):
        super().__init__()
        from ....models.vision.clip_vit_model import CLIPViTModel
        from ....transformer.norms import FusedLayerNorm

        self.config, self.patch_dim, self.img_h, self.img_w = config, patch_dim, img_h, img_w
        self.backbone = CLIPViTModel(config, None, ln_pre_impl=None, ln_post_impl=FusedLayerNorm, add_class_token=False,
                                     patch_dim=patch_dim, img_h=img_h, img_w=img_w)
        self.linear_decoder = nn.Linear(config.hidden_size, 3 * patch_dim * patch_dim, dtype=config.params_dtype)
        self.share_embeddings_and_output_weights = False

    def set_input_tensor(self, t):
        self.backbone.set_input_tensor(t)

    def forward(self, images: torch.Tensor) -> torch.Tensor:
        hidden = self.backbone(images)                                   # [b, patches, h]
        p, gh, gw = self.patch_dim, self.img_h // self.patch_dim, self.img_w // self.patch_dim
        y = self.linear_decoder(hidden).view(-1, gh, gw, p, p, 3)
        return y.permute(0, 5, 1, 3, 2, 4).reshape(-1, 3, gh * p, gw * p)


class _MLPProj(nn.Module):
    def __init__(self, cin, dim):
        super().__init__()
        self.proj = nn.Linear(cin, dim)

    def forward(self, x):
        return self.proj(x.flatten(2).transpose(1, 2))


class MitInpaintingModel(nn.Module):
    def __init__(self, config=None, embedding_dim=256, variant="b0", pre_process=True, post_process=True):
        super().__init__()
        from . import mit_backbone

        self.config = config
        self.backbone = getattr(mit_backbone, f"mit_{variant}")()
        dims = self.backbone.embed_dims
        self.projs = nn.ModuleList([_MLPProj(c, embedding_dim) for c in dims])
        self.fuse = nn.Sequential(nn.Conv2d(4 * embedding_dim, embedding_dim, 1, bias=False), nn.BatchNorm2d(embedding_dim), nn.ReLU())
        self.color = nn.Conv2d(embedding_dim, 3 * 16, 1)                 # 4x4 pixel shuffle back to full resolution
        self.share_embeddings_and_output_weights = False

    def set_input_tensor(self, t):
        pass

    def forward(self, images):
        feats = self.backbone(images)
        H, W = feats[0].shape[2:]
        ups = []
        for f, pr in zip(feats, self.projs):
            y = pr(f).transpose(1, 2).reshape(f.shape[0], -1, f.shape[2], f.shape[3])
            ups.append(F.interpolate(y, size=(H, W), mode="bilinear", align_corners=False))
        y = self.color(self.fuse(torch.cat(ups[::-1], dim=1)))
        return F.pixel_shuffle(y, 4)


def masked_reconstruction_loss(pred, images, masks):
    """L1 + (1 - PSNR-proxy) on the masked pixels only; ``masks`` is 1 where the input was erased."""
    m = masks.to(pred.dtype)
    denom = m.sum().clamp_min(1.0) * pred.shape[1] / m.shape[1]
    l1 = ((pred.float() - images.float()).abs() * m).sum() / denom
    mse = (((pred.float() - images.float()) ** 2) * m).sum() / denom
    psnr = -10.0 * torch.log10(mse.clamp_min(1e-10))
    return l1, psnr
