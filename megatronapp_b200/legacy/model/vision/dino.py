"""DINO self-distillation (reference ``legacy/model/vision/dino.py``, 291 lines + ``pretrain_vision_dino.py``): a student and
a momentum (EMA) teacher share an architecture; the student matches the centred, sharpened teacher distribution of the
*global* crops on every *other* crop.  Backbones: the framework ViT (``CLIPViTModel``), Swin or MiT."""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F


class DINOHead(nn.Module):
    def __init__(self, in_dim, out_dim, hidden=2048, bottleneck=256, nlayers=3, norm_last_layer=True):
        super().__init__()
        layers, d = [], in_dim
        for _ in range(max(nlayers - 1, 0)):
            layers += [nn.Linear(d, hidden), nn.GELU()]
            d = hidden
        layers.append(nn.Linear(d, bottleneck))
        self.mlp = nn.Sequential(*layers)
        for m in self.mlp:
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)
        self.last = nn.utils.parametrizations.weight_norm(nn.Linear(bottleneck, out_dim, bias=False))
        self.last.parametrizations.weight.original0.data.fill_(1.0)
        if norm_last_layer:
            self.last.parametrizations.weight.original0.requires_grad = False

    def forward(self, x):
        return self.last(F.normalize(self.mlp(x), dim=-1, p=2))


class DINOLoss(nn.Module):
    def __init__(self, out_dim, ncrops, warmup_teacher_temp=0.04, teacher_temp=0.07, warmup_epochs=30, nepochs=100,
                 student_temp=0.1, center_momentum=0.9):
        super().__init__()
        self.student_temp, self.center_momentum, self.ncrops = student_temp, center_momentum, ncrops
        self.register_buffer("center", torch.zeros(1, out_dim))
        w = min(warmup_epochs, nepochs)
        self.teacher_temp_schedule = torch.cat([torch.linspace(warmup_teacher_temp, teacher_temp, w),
                                                torch.full((max(nepochs - w, 0),), teacher_temp)])

    def forward(self, student_out: torch.Tensor, teacher_out: torch.Tensor, epoch: int) -> torch.Tensor:
        s = (student_out / self.student_temp).chunk(self.ncrops)
        temp = float(self.teacher_temp_schedule[min(epoch, len(self.teacher_temp_schedule) - 1)])
        t = F.softmax((teacher_out - self.center) / temp, dim=-1).detach().chunk(2)
        total, n = 0.0, 0
        for iq, q in enumerate(t):
            for v in range(len(s)):
                if v == iq:
                    continue                                  # same view: skipped
                total = total + torch.sum(-q * F.log_softmax(s[v], dim=-1), dim=-1).mean()
                n += 1
        self.update_center(teacher_out)
        return total / n

    @torch.no_grad()
    def update_center(self, teacher_out):
        c = teacher_out.sum(0, keepdim=True)
        n = teacher_out.shape[0]
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(c)
            n *= torch.distributed.get_world_size()
        self.center = self.center * self.center_momentum + (c / n) * (1 - self.center_momentum)


def cosine_scheduler(base, final, epochs, niter_per_ep):
    it = torch.arange(epochs * niter_per_ep, dtype=torch.float32)
    return final + 0.5 * (base - final) * (1 + torch.cos(math.pi * it / max(1, len(it))))


class _Backbone(nn.Module):
    """image batch -> pooled feature, for any of the supported trunks"""

    def __init__(self, kind, config, img_size, patch_dim, swin_type="tiny", is_teacher=False):
        super().__init__()
        self.kind = kind
        if kind == "vit":
            from ....models.vision.clip_vit_model import CLIPViTModel
            from ....transformer.norms import FusedLayerNorm

            self.net = CLIPViTModel(config, None, ln_pre_impl=None, ln_post_impl=FusedLayerNorm, add_class_token=True,
                                    patch_dim=patch_dim, img_h=img_size, img_w=img_size)
            self.dim = config.hidden_size
        elif kind == "swin":
            from .swin_backbone import get_swin

            self.net = get_swin(img_size=img_size, output_avg=True, kind=swin_type, is_teacher=is_teacher)
            self.dim = self.net.num_features
        else:
            from .mit_backbone import mit_b0

            self.net = mit_b0(output_avg=True)
            self.dim = self.net.embed_dims[-1]

    def forward(self, x):
        y = self.net(x)
        return y[:, 0, :] if self.kind == "vit" else y


class DINOPretrainModel(nn.Module):
    """forward(list of crops: 2 global first, then local) -> (student logits for all crops, teacher logits for the 2 global
    crops); ``update_momentum(it)`` is the EMA teacher update after the optimizer step."""

    def __init__(self, config, out_dim=65536, img_size=224, local_img_size=96, patch_dim=16, backbone="vit",
                 norm_last_layer=True, momentum_teacher=0.996, epochs=100, iters_per_epoch=1000, head_hidden=2048, bottleneck=256,
                 swin_type="tiny"):
        super().__init__()
        self.config = config
        self.student = _Backbone(backbone, config, img_size, patch_dim, swin_type)
        self.teacher = _Backbone(backbone, config, img_size, patch_dim, swin_type, is_teacher=True)
        self.student_head = DINOHead(self.student.dim, out_dim, head_hidden, bottleneck, norm_last_layer=norm_last_layer)
        self.teacher_head = DINOHead(self.teacher.dim, out_dim, head_hidden, bottleneck)
        self.teacher.load_state_dict(self.student.state_dict())
        self.teacher_head.load_state_dict(self.student_head.state_dict(), strict=False)
        for p in list(self.teacher.parameters()) + list(self.teacher_head.parameters()):
            p.requires_grad = False
        self.momentum_schedule = cosine_scheduler(momentum_teacher, 1.0, epochs, iters_per_epoch)
        self.img_size, self.local_img_size = img_size, local_img_size
        self.share_embeddings_and_output_weights = False

    def set_input_tensor(self, t):
        pass

    def _multi_res(self, net, head, crops: List[torch.Tensor]):
        # crops of one resolution go through the trunk as one batch (ViT position embeddings are resolution specific)
        outs, i = [], 0
        while i < len(crops):
            j = i
            while j < len(crops) and crops[j].shape[-1] == crops[i].shape[-1]:
                j += 1
            x = torch.cat(crops[i:j])
            if x.shape[-1] != self.img_size and getattr(net, "kind", "") == "vit":
                x = F.interpolate(x, size=(self.img_size, self.img_size), mode="bilinear", align_corners=False)
            outs.append(net(x))
            i = j
        return head(torch.cat(outs))

    def forward(self, crops: List[torch.Tensor]):
        student = self._multi_res(self.student, self.student_head, crops)
        with torch.no_grad():
            teacher = self._multi_res(self.teacher, self.teacher_head, crops[:2])
        return student, teacher

    @torch.no_grad()
    def update_momentum(self, iteration: int):
        m = float(self.momentum_schedule[min(iteration, len(self.momentum_schedule) - 1)])
        for ps_, pt in zip(list(self.student.parameters()) + list(self.student_head.parameters()),
                           list(self.teacher.parameters()) + list(self.teacher_head.parameters())):
            if ps_.shape == pt.shape:
                pt.mul_(m).add_(ps_.detach().to(pt.dtype), alpha=1 - m)
