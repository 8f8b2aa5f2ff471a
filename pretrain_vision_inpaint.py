"""Image in-painting pre-training (reference ``pretrain_vision_inpaint.py``): random (or row) patch masks erase part of
the input; the model (ViT + linear decoder, or MiT + all-MLP decoder) reconstructs it; L1 on the masked pixels, PSNR
reported."""
import os

if not os.path.exists("/dev/nvidia0"):
    os.environ.setdefault("OMP_NUM_THREADS", "2")

from functools import partial

import torch

from megatronapp_b200.legacy.model.vision.inpainting import MitInpaintingModel, VitInpaintingModel, masked_reconstruction_loss
from megatronapp_b200.parallel import state as mpu
from megatronapp_b200.training.arguments import core_transformer_config_from_args
from megatronapp_b200.training.global_vars import get_args
from megatronapp_b200.training.training import pretrain
from megatronapp_b200.training.utils import print_rank_0


def model_provider(pre_process=True, post_process=True):
    a = get_args()
    cfg = core_transformer_config_from_args(a)
    print_rank_0(f"building {a.vision_backbone_type} in-painting model ...")
    if a.vision_backbone_type == "vit":
        return VitInpaintingModel(cfg, a.img_h, a.img_w, a.patch_dim, pre_process=pre_process, post_process=post_process)
    if a.vision_backbone_type == "mit":
        return MitInpaintingModel(cfg, pre_process=pre_process, post_process=post_process)
    raise ValueError(f"in-painting supports the vit and mit backbones, not {a.vision_backbone_type}")


class MockInpaintDataset(torch.utils.data.Dataset):
    """Smooth synthetic images (low-frequency colour fields -- reconstructable from context) with patch masks:
    ``random``: each patch is erased with probability ``mask_factor / 4`` ; ``row``: a contiguous band of patch rows."""

    def __init__(self, n, h, w, patch, mask_type, mask_factor, seed):
        self.n, self.h, self.w, self.p, self.mask_type, self.mask_factor, self.seed = n, h, w, patch, mask_type, mask_factor, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        coarse = torch.rand(1, 3, 3, 3, generator=g)
        img = torch.nn.functional.interpolate(coarse, size=(self.h, self.w), mode="bilinear", align_corners=True)[0]
        gh, gw = self.h // self.p, self.w // self.p
        if self.mask_type == "row":
            k = max(1, int(gh * min(0.5, self.mask_factor / 4)))
            r0 = int(torch.randint(0, gh - k + 1, (1,), generator=g))
            pm = torch.zeros(gh, gw)
            pm[r0:r0 + k] = 1
        else:
            pm = (torch.rand(gh, gw, generator=g) < min(0.9, self.mask_factor / 4)).float()
            if pm.sum() == 0:
                pm[int(torch.randint(0, gh, (1,), generator=g)), int(torch.randint(0, gw, (1,), generator=g))] = 1
        mask = pm.repeat_interleave(self.p, 0).repeat_interleave(self.p, 1)[None]
        return {"image": img, "mask": mask}


class _MaskedView(torch.utils.data.Dataset):
    def __init__(self, ds, n):
        self.ds, self.n = ds, max(int(n), len(ds))

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        (img, mask), _ = self.ds[i % len(self.ds)]
        return {"image": img, "mask": mask}


def train_valid_test_datasets_provider(num_samples):
    a = get_args()
    if a.data_path:
        from megatronapp_b200.legacy.data.vit_dataset import build_train_valid_datasets

        tr, va = build_train_valid_datasets(a.data_path if len(a.data_path) > 1 else a.data_path[0], (a.img_h, a.img_w), "inpaint",
                                            patch_dim=a.patch_dim, mask_factor=a.mask_factor, mask_type=a.mask_type,
                                            classes_fraction=a.classes_fraction, data_per_class_fraction=a.data_per_class_fraction)
        return _MaskedView(tr, num_samples[0]), _MaskedView(va, num_samples[1]), _MaskedView(va, num_samples[2])
    mk = lambda n, s: MockInpaintDataset(max(int(n), 1), a.img_h, a.img_w, a.patch_dim, a.mask_type, a.mask_factor, s)  # noqa: E731
    return mk(num_samples[0], 1), mk(num_samples[1], 2), mk(num_samples[2], 3)


def loss_func(images, masks, output):
    l1, psnr = masked_reconstruction_loss(output, images, masks)
    rep = torch.stack([l1.detach(), psnr.detach()])
    grp = mpu.get_data_parallel_group()
    if grp is not None and not isinstance(grp, mpu._SingleGroup) and torch.distributed.get_world_size(grp) > 1:
        torch.distributed.all_reduce(rep, group=grp)
        rep /= torch.distributed.get_world_size(grp)
    return l1, {"loss": rep[0], "psnr": rep[1]}


def forward_step(data_iterator, model):
    p = next(model.parameters())
    d = next(data_iterator)
    images, masks = d["image"].to(p.device), d["mask"].to(p.device)
    out = model((images * (1 - masks)).to(p.dtype))
    return out, partial(loss_func, images, masks)


if __name__ == "__main__":
    train_valid_test_datasets_provider.is_distributed = True
    pretrain(train_valid_test_datasets_provider, model_provider, "encoder_or_decoder", forward_step,
             args_defaults={"tokenizer_type": "NullTokenizer", "vocab_size": 8, "vision_pretraining": True,
                            "vision_pretraining_type": "inpaint"})
