"""Vision pre-training datasets (reference ``legacy/data/vit_dataset.py``): image folders under ``<data_path>/train`` and
``<data_path>/val`` with the transform of the chosen task --

* ``classify``: random resized crop + flip + colour jitter + AutoAugment (train) / resize + centre crop (eval);
* ``inpaint``:  the same image pipeline plus a random patch mask (``random``: i.i.d. patches, ``row``: horizontal band);
* ``dino``:     multi-crop -- two global views and ``n`` local views with the DINO colour / blur / solarise schedule."""
from __future__ import annotations

import os
import random
from typing import Tuple

import numpy as np
import torch
import torchvision.transforms as T
from PIL import ImageFilter, ImageOps

from .autoaugment import ImageNetPolicy
from .image_folder import ImageFolder

_MEAN, _STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


class GaussianBlur:
    def __init__(self, p=0.5, radius_min=0.1, radius_max=2.0):
        self.p, self.lo, self.hi = p, radius_min, radius_max

    def __call__(self, img):
        return img.filter(ImageFilter.GaussianBlur(radius=random.uniform(self.lo, self.hi))) if random.random() <= self.p else img


class Solarization:
    def __init__(self, p):
        self.p = p

    def __call__(self, img):
        return ImageOps.solarize(img) if random.random() < self.p else img


class ClassificationTransform:
    def __init__(self, image_size, train=True, dtype=torch.float32):
        size = image_size if isinstance(image_size, (tuple, list)) else (image_size, image_size)
        if train:
            steps = [T.RandomResizedCrop(size), T.RandomHorizontalFlip(), T.ColorJitter(0.4, 0.4, 0.4, 0.1), ImageNetPolicy()]
        else:
            steps = [T.Resize(size), T.CenterCrop(size)]
        self.transform = T.Compose(steps + [T.ToTensor(), T.Normalize(_MEAN, _STD), T.ConvertImageDtype(dtype)])

    def __call__(self, img):
        return self.transform(img)


class InpaintingTransform:
    def __init__(self, image_size, patch_dim, mask_factor=1.0, mask_type="random", train=True, dtype=torch.float32):
        size = image_size if isinstance(image_size, (tuple, list)) else (image_size, image_size)
        self.size, self.patch, self.mask_factor, self.mask_type, self.train = size, patch_dim, mask_factor, mask_type, train
        steps = [T.RandomResizedCrop(size), T.RandomHorizontalFlip(), T.ColorJitter(0.4, 0.4, 0.4, 0.1), ImageNetPolicy()] if train \
            else [T.Resize(size, interpolation=2), T.CenterCrop(size)]
        self.transform = T.Compose(steps + [T.ToTensor(), T.ConvertImageDtype(dtype)])

    def gen_mask(self) -> torch.Tensor:
        gh, gw = self.size[0] // self.patch, self.size[1] // self.patch
        frac = min(0.9, self.mask_factor / 4)
        rng = np.random if self.train else np.random.RandomState(0)            # evaluation uses one fixed mask
        if self.mask_type == "row":
            k = max(1, int(gh * frac))
            r0 = int(rng.randint(0, gh - k + 1))
            pm = np.zeros((gh, gw), dtype=np.float32)
            pm[r0:r0 + k] = 1
        else:
            pm = (rng.rand(gh, gw) < frac).astype(np.float32)
            if pm.sum() == 0:
                pm[int(rng.randint(0, gh)), int(rng.randint(0, gw))] = 1
        return torch.from_numpy(np.kron(pm, np.ones((self.patch, self.patch), dtype=np.float32)))[None]

    def __call__(self, img):
        return self.transform(img), self.gen_mask()


class DinoTransform:
    def __init__(self, image_size, local_size=96, local_crops_number=10, train=True, dtype=torch.float32,
                 global_crops_scale=(0.4, 1.0), local_crops_scale=(0.05, 0.4)):
        self.train, self.n_local = train, local_crops_number
        flip_color = T.Compose([T.RandomHorizontalFlip(p=0.5), T.RandomApply([T.ColorJitter(0.4, 0.4, 0.2, 0.1)], p=0.8),
                                T.RandomGrayscale(p=0.2)])
        norm = T.Compose([T.ToTensor(), T.Normalize(_MEAN, _STD), T.ConvertImageDtype(dtype)])
        crop = lambda size, scale: T.RandomResizedCrop(size, scale=scale, interpolation=3)  # noqa: E731  bicubic
        self.global1 = T.Compose([crop(image_size, global_crops_scale), flip_color, GaussianBlur(1.0), norm])
        self.global2 = T.Compose([crop(image_size, global_crops_scale), flip_color, GaussianBlur(0.1), Solarization(0.2), norm])
        self.local = T.Compose([crop(local_size, local_crops_scale), flip_color, GaussianBlur(0.5), norm])
        self.eval = T.Compose([T.Resize(image_size, interpolation=3), T.CenterCrop(image_size), norm])

    def __call__(self, img):
        if not self.train:
            return self.eval(img)
        return [self.global1(img), self.global2(img)] + [self.local(img) for _ in range(self.n_local)]


def build_train_valid_datasets(data_path, image_size=224, vision_pretraining_type="classify", patch_dim=16, mask_factor=1.0,
                               mask_type="random", dino_local_img_size=96, dino_local_crops_number=10, classes_fraction=1.0,
                               data_per_class_fraction=1.0, dtype=torch.float32) -> Tuple[ImageFolder, ImageFolder]:
    """``data_path``: a directory with ``train`` and ``val`` sub-directories, or ``[train_dir, val_dir]``."""
    if isinstance(data_path, (list, tuple)):
        train_dir, val_dir = data_path[0], data_path[1 if len(data_path) > 1 else 0]
    else:
        train_dir, val_dir = os.path.join(data_path, "train"), os.path.join(data_path, "val")
        if not os.path.isdir(val_dir):
            val_dir = train_dir
    if vision_pretraining_type == "classify":
        tr, va = ClassificationTransform(image_size, True, dtype), ClassificationTransform(image_size, False, dtype)
    elif vision_pretraining_type == "inpaint":
        tr = InpaintingTransform(image_size, patch_dim, mask_factor, mask_type, True, dtype)
        va = InpaintingTransform(image_size, patch_dim, mask_factor, mask_type, False, dtype)
    elif vision_pretraining_type == "dino":
        tr = DinoTransform(image_size, dino_local_img_size, dino_local_crops_number, True, dtype)
        va = ClassificationTransform(image_size, False, dtype)
    else:
        raise ValueError(f"{vision_pretraining_type} vision pretraining type is not supported")
    return (ImageFolder(train_dir, tr, classes_fraction, data_per_class_fraction), ImageFolder(val_dir, va))
