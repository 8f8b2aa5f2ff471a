"""Image-classification finetuning of the ViT backbone (reference ``tasks/vision/classification/classification.py`` +
``tasks/vision/finetune_utils.py``): an image-folder dataset (one sub-directory per class holding ``.npy`` / ``.pt``
``[3, H, W]`` tensors -- no torchvision dependency), epoch loop, top-1 accuracy per epoch, optional warm start from a
pre-trained backbone via ``--pretrained-checkpoint``."""
from __future__ import annotations

import os
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F


class ImageFolderTensors(torch.utils.data.Dataset):
    def __init__(self, root: str, img_h: int, img_w: int, classes: List[str] = None):
        self.classes = classes or sorted(d for d in os.listdir(root) if os.path.isdir(os.path.join(root, d)))
        self.items = [(os.path.join(root, c, f), i) for i, c in enumerate(self.classes)
                      for f in sorted(os.listdir(os.path.join(root, c))) if f.endswith((".npy", ".pt"))]
        self.size = (img_h, img_w)

    def __len__(self):
        return len(self.items)

    def __getitem__(self, idx) -> Dict[str, torch.Tensor]:
        path, label = self.items[idx]
        img = torch.load(path) if path.endswith(".pt") else torch.from_numpy(np.load(path))
        img = img.float()
        if tuple(img.shape[-2:]) != self.size:
            img = F.interpolate(img.unsqueeze(0), size=self.size, mode="bilinear", align_corners=False)[0]
        return {"image": img, "label": torch.tensor(label)}


def vision_forward_step(batch, model):
    dev = next(model.parameters()).device
    logits = model(batch["image"].to(dev))
    labels = batch["label"].long().to(dev)

    def loss_fn(out):
        loss = F.cross_entropy(out.float(), labels)
        return loss, {"loss": loss.detach()}

    return logits, loss_fn


@torch.no_grad()
def top1(model, loader) -> Dict[str, float]:
    import torch.distributed as dist

    from megatronapp_b200.parallel import state as ps

    was = model.training
    model.eval()
    dev = next(model.parameters()).device
    mp = ps.model_parallel_is_initialized()
    dp_rank, dp_size = (ps.get_data_parallel_rank(), ps.get_data_parallel_world_size()) if mp else (0, 1)
    t = torch.zeros(2, dtype=torch.float64, device=dev)
    for i, b in enumerate(loader):
        if i % dp_size == dp_rank:
            pred = model(b["image"].to(dev)).argmax(-1)
            t[0] += (pred == b["label"].to(dev)).sum()
            t[1] += pred.numel()
    if mp and dist.is_initialized():
        g = ps.get_data_parallel_group()
        if g is not None and not isinstance(g, ps._SingleGroup):
            dist.all_reduce(t, group=g)
    model.train(was)
    return {"correct": int(t[0]), "total": int(t[1]), "accuracy": float(t[0] / t[1].clamp_min(1))}


def main():
    from megatronapp_b200.models.vision.vit_classification import VitClassificationModel
    from megatronapp_b200.training import checkpointing
    from megatronapp_b200.training.arguments import core_transformer_config_from_args
    from megatronapp_b200.training.global_vars import get_args
    from megatronapp_b200.training.utils import print_rank_0
    from tasks.finetune_utils import finetune

    args = get_args()
    cfg = core_transformer_config_from_args(args)
    model = VitClassificationModel(cfg, args.num_classes, finetune=True, img_h=args.img_h, img_w=args.img_w,
                                   patch_dim=args.patch_dim)
    if args.pretrained_checkpoint:
        args.load, args.finetune, args.no_load_optim, args.no_load_rng = args.pretrained_checkpoint, True, True, True
        checkpointing.load_checkpoint([model], None, None, strict=False)
    train = ImageFolderTensors(args.train_data[0], args.img_h, args.img_w)
    valid = ImageFolderTensors(args.valid_data[0], args.img_h, args.img_w, train.classes)
    vloader = torch.utils.data.DataLoader(valid, batch_size=args.micro_batch_size, shuffle=False)
    hist = finetune(model, train, valid, epochs=args.epochs, batch_size=args.micro_batch_size, lr=args.lr,
                    forward_step=vision_forward_step, evaluate=lambda m, _l: top1(m, vloader))
    for h in hist:
        print_rank_0(f" >> |epoch: {h['epoch']}| correct / total: {h['correct']} / {h['total']} = {100 * h['accuracy']:.4f} %")
    return hist
