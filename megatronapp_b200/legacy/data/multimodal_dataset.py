"""(caption, image) dataset over the multimodal ``.bin/.idx`` written by ``tools/preprocess_mmdata.py`` (reference
``legacy/data/multimodal_dataset.py``): every document = a mode-0 sequence (caption token ids) followed by a mode-1 sequence
(the encoded image file, int32 words with the pad count in front).  Images are decoded with PIL and go through a CLIP-style
random-resized-crop + normalise transform."""
from __future__ import annotations

import io

import numpy as np
import torch
from PIL import Image
from torchvision.transforms import Compose, InterpolationMode, Normalize, RandomResizedCrop, ToTensor

from ...data.indexed_dataset import IndexedDataset

CLIP_MEAN, CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)


def _transform(img_h: int, img_w: int):
    return Compose([RandomResizedCrop((img_h, img_w), scale=(0.5, 1.0), interpolation=InterpolationMode.BICUBIC),
                    lambda im: im.convert("RGB"), ToTensor(), Normalize(CLIP_MEAN, CLIP_STD)])


class MultiModalDataset(torch.utils.data.Dataset):
    def __init__(self, name: str, data_prefix: str, indexed_dataset: IndexedDataset = None, num_samples: int = None,
                 seq_length: int = None, seed: int = 1234, img_h: int = 224, img_w: int = 224, flatten_image: bool = False):
        self.name = name
        self.ds = indexed_dataset or IndexedDataset(data_prefix, multimodal=True)
        self.doc_idx = np.asarray(self.ds.document_indices)
        self.n_docs = len(self.doc_idx) - 1
        self.num_samples = num_samples or self.n_docs
        self.visual_transform = _transform(img_h, img_w)
        self.flatten_image = flatten_image

    def __len__(self):
        return self.num_samples

    def __getitem__(self, idx):
        first = int(self.doc_idx[idx % self.n_docs])
        modes = self.ds.index.sequence_modes
        if int(modes[first]) != 0 or int(modes[first + 1]) != 1:
            raise ValueError(f"document {idx}: expected (text, image) sequences, got modes {modes[first]}, {modes[first + 1]}")
        text = np.asarray(self.ds.get(first), dtype=np.int64)
        words = np.asarray(self.ds.get(first + 1), dtype=np.int32)
        pad = int(words[0])
        raw = words[1:].tobytes(order="C")
        raw = raw[: len(raw) - pad] if pad else raw
        img = self.visual_transform(Image.open(io.BytesIO(raw)))
        return {"text": text, "img": (img.reshape(-1) if self.flatten_image else img).numpy().astype(np.float32)}
