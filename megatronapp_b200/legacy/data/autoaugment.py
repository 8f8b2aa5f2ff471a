"""AutoAugment for ImageNet (reference ``legacy/data/autoaugment.py``; Cubuk et al. 2019): 25 learned sub-policies, each two
(operation, probability, magnitude level) steps on a PIL image; one sub-policy is drawn per image."""
from __future__ import annotations

import random

import numpy as np
from PIL import Image, ImageEnhance, ImageOps

_LEVELS = 10
_RANGES = {
    "shearX": np.linspace(0, 0.3, _LEVELS), "shearY": np.linspace(0, 0.3, _LEVELS),
    "translateX": np.linspace(0, 150 / 331, _LEVELS), "translateY": np.linspace(0, 150 / 331, _LEVELS),
    "rotate": np.linspace(0, 30, _LEVELS), "color": np.linspace(0.0, 0.9, _LEVELS),
    "posterize": np.round(np.linspace(8, 4, _LEVELS), 0).astype(int), "solarize": np.linspace(256, 0, _LEVELS),
    "contrast": np.linspace(0.0, 0.9, _LEVELS), "sharpness": np.linspace(0.0, 0.9, _LEVELS),
    "brightness": np.linspace(0.0, 0.9, _LEVELS), "autocontrast": [0] * _LEVELS, "equalize": [0] * _LEVELS, "invert": [0] * _LEVELS,
}


def _sign(rng):
    return rng.choice((-1, 1))


def _rotate_with_fill(img, deg, fill):
    rot = img.convert("RGBA").rotate(deg)
    return Image.composite(rot, Image.new("RGBA", rot.size, fill + (255,)), rot).convert(img.mode)


def _apply(op, img, mag, fill, rng):
    w, h = img.size
    if op == "shearX":
        return img.transform(img.size, Image.AFFINE, (1, mag * _sign(rng), 0, 0, 1, 0), Image.BICUBIC, fillcolor=fill)
    if op == "shearY":
        return img.transform(img.size, Image.AFFINE, (1, 0, 0, mag * _sign(rng), 1, 0), Image.BICUBIC, fillcolor=fill)
    if op == "translateX":
        return img.transform(img.size, Image.AFFINE, (1, 0, mag * w * _sign(rng), 0, 1, 0), fillcolor=fill)
    if op == "translateY":
        return img.transform(img.size, Image.AFFINE, (1, 0, 0, 0, 1, mag * h * _sign(rng)), fillcolor=fill)
    if op == "rotate":
        return _rotate_with_fill(img, mag * _sign(rng), fill)
    if op == "color":
        return ImageEnhance.Color(img).enhance(1 + mag * _sign(rng))
    if op == "posterize":
        return ImageOps.posterize(img, int(mag))
    if op == "solarize":
        return ImageOps.solarize(img, mag)
    if op == "contrast":
        return ImageEnhance.Contrast(img).enhance(1 + mag * _sign(rng))
    if op == "sharpness":
        return ImageEnhance.Sharpness(img).enhance(1 + mag * _sign(rng))
    if op == "brightness":
        return ImageEnhance.Brightness(img).enhance(1 + mag * _sign(rng))
    if op == "autocontrast":
        return ImageOps.autocontrast(img)
    if op == "equalize":
        return ImageOps.equalize(img)
    if op == "invert":
        return ImageOps.invert(img)
    raise ValueError(op)


class SubPolicy:
    def __init__(self, p1, op1, level1, p2, op2, level2, fillcolor=(128, 128, 128)):
        self.steps = ((p1, op1, _RANGES[op1][level1]), (p2, op2, _RANGES[op2][level2]))
        self.fill = fillcolor

    def __call__(self, img, rng=random):
        for p, op, mag in self.steps:
            if rng.random() < p:
                img = _apply(op, img, mag, self.fill, rng)
        return img


# the ImageNet policy table of the paper (probability, operation, level) x 2
_IMAGENET = [
    (0.4, "posterize", 8, 0.6, "rotate", 9), (0.6, "solarize", 5, 0.6, "autocontrast", 5), (0.8, "equalize", 8, 0.6, "equalize", 3),
    (0.6, "posterize", 7, 0.6, "posterize", 6), (0.4, "equalize", 7, 0.2, "solarize", 4), (0.4, "equalize", 4, 0.8, "rotate", 8),
    (0.6, "solarize", 3, 0.6, "equalize", 7), (0.8, "posterize", 5, 1.0, "equalize", 2), (0.2, "rotate", 3, 0.6, "solarize", 8),
    (0.6, "equalize", 8, 0.4, "posterize", 6), (0.8, "rotate", 8, 0.4, "color", 0), (0.4, "rotate", 9, 0.6, "equalize", 2),
    (0.0, "equalize", 7, 0.8, "equalize", 8), (0.6, "invert", 4, 1.0, "equalize", 8), (0.6, "color", 4, 1.0, "contrast", 8),
    (0.8, "rotate", 8, 1.0, "color", 2), (0.8, "color", 8, 0.8, "solarize", 7), (0.4, "sharpness", 7, 0.6, "invert", 8),
    (0.6, "shearX", 5, 1.0, "equalize", 9), (0.4, "color", 0, 0.6, "equalize", 3), (0.4, "equalize", 7, 0.2, "solarize", 4),
    (0.6, "solarize", 5, 0.6, "autocontrast", 5), (0.6, "invert", 4, 1.0, "equalize", 8), (0.6, "color", 4, 1.0, "contrast", 8),
    (0.8, "equalize", 8, 0.6, "equalize", 3),
]


class ImageNetPolicy:
    def __init__(self, fillcolor=(128, 128, 128)):
        self.policies = [SubPolicy(*row, fillcolor=fillcolor) for row in _IMAGENET]

    def __call__(self, img):
        return random.choice(self.policies)(img)

    def __repr__(self):
        return "ImageNetPolicy"
