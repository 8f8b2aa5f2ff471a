"""Image-folder dataset (reference ``legacy/data/image_folder.py``): ``root/<class>/<image>``; ``classes_fraction`` keeps the
first fraction of the (sorted) classes, ``data_per_class_fraction`` the first fraction of every class's files -- the knobs of
``--classes-fraction`` / ``--data-per-class-fraction`` for low-data experiments."""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
from PIL import Image

IMG_EXTENSIONS = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".pgm", ".tif", ".tiff", ".webp", ".npy")


def pil_loader(path: str) -> Image.Image:
    if path.endswith(".npy"):
        a = np.load(path)
        if a.ndim == 3 and a.shape[0] in (1, 3):
            a = a.transpose(1, 2, 0)
        if a.dtype != np.uint8:
            a = (np.clip(a, 0.0, 1.0) * 255).astype(np.uint8)
        return Image.fromarray(a.squeeze()).convert("RGB")
    with open(path, "rb") as f:
        return Image.open(f).convert("RGB")


class ImageFolder(torch.utils.data.Dataset):
    def __init__(self, root: str, transform: Optional[Callable] = None, classes_fraction: float = 1.0,
                 data_per_class_fraction: float = 1.0, loader: Callable = pil_loader):
        classes = sorted(d.name for d in os.scandir(root) if d.is_dir())
        if not classes:
            raise FileNotFoundError(f"no class folders under {root}")
        classes = classes[: max(1, int(len(classes) * classes_fraction))]
        self.classes, self.class_to_idx = classes, {c: i for i, c in enumerate(classes)}
        self.samples: List[Tuple[str, int]] = []
        for c in classes:
            files = sorted(os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(root, c)) for f in fs
                           if f.lower().endswith(IMG_EXTENSIONS))
            files = files[: max(1, int(len(files) * data_per_class_fraction))] if files else []
            self.samples += [(f, self.class_to_idx[c]) for f in files]
        if not self.samples:
            raise FileNotFoundError(f"no images with extensions {IMG_EXTENSIONS} under {root}")
        self.root, self.transform, self.loader = root, transform, loader
        self.total = len(self.samples)

    def __len__(self):
        return self.total

    def __getitem__(self, i):
        path, target = self.samples[i % self.total]
        img = self.loader(path)
        return (self.transform(img) if self.transform is not None else img), target
