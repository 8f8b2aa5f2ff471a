"""Pretrain a ViT classifier (reference ``pretrain_vision_classify.py``); mock images unless ``--data-path`` points at
an ImageFolder-style directory readable by torchvision (not part of this image)."""
import os

if not os.path.exists("/dev/nvidia0"):
    os.environ.setdefault("OMP_NUM_THREADS", "2")

from functools import partial

import torch
import torch.nn.functional as F

from megatronapp_b200.models.vision.vit_classification import VitClassificationModel
from megatronapp_b200.parallel import state as mpu
from megatronapp_b200.training.arguments import core_transformer_config_from_args
from megatronapp_b200.training.global_vars import get_args
from megatronapp_b200.training.training import pretrain
from megatronapp_b200.training.utils import print_rank_0


def add_vision_args(parser):
    g = parser.add_argument_group("vision")
    g.add_argument("--num-classes", type=int, default=1000)
    g.add_argument("--img-h", type=int, default=224)
    g.add_argument("--img-w", type=int, default=224)
    g.add_argument("--patch-dim", type=int, default=16)
    return parser


def model_provider(pre_process=True, post_process=True):
    args = get_args()
    print_rank_0("building ViT classification model ...")
    cfg = core_transformer_config_from_args(args)
    return VitClassificationModel(cfg, args.num_classes, args.img_h, args.img_w, args.patch_dim,
                                  pre_process=pre_process, post_process=post_process)


class MockImageDataset(torch.utils.data.Dataset):
    def __init__(self, n, h, w, classes, seed):
        self.n, self.h, self.w, self.classes, self.seed = n, h, w, classes, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        label = int(torch.randint(0, self.classes, (1,), generator=g))
        img = torch.rand(3, self.h, self.w, generator=g) * 0.1
        img[label % 3] += 0.5 + 0.4 * (label / max(1, self.classes - 1))      # class-dependent colour shift: learnable signal
        return {"image": img, "label": torch.tensor(label)}


class _DictView(torch.utils.data.Dataset):
    """(image, label) folder samples as the dict the forward step reads"""

    def __init__(self, ds, n):
        self.ds, self.n = ds, max(int(n), len(ds))

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        img, label = self.ds[i % len(self.ds)]
        return {"image": img, "label": torch.tensor(label)}


def train_valid_test_datasets_provider(num_samples):
    a = get_args()
    if a.data_path:
        # <data-path>/train/<class>/*.jpg and <data-path>/val/... (reference legacy/data/vit_dataset.py)
        from megatronapp_b200.legacy.data.vit_dataset import build_train_valid_datasets

        tr, va = build_train_valid_datasets(a.data_path if len(a.data_path) > 1 else a.data_path[0], (a.img_h, a.img_w), "classify",
                                            classes_fraction=a.classes_fraction, data_per_class_fraction=a.data_per_class_fraction)
        return _DictView(tr, num_samples[0]), _DictView(va, num_samples[1]), _DictView(va, num_samples[2])
    mk = lambda n, s: MockImageDataset(max(int(n), 1), a.img_h, a.img_w, a.num_classes, s)  # noqa: E731
    return mk(num_samples[0], 1), mk(num_samples[1], 2), mk(num_samples[2], 3)


def loss_func(labels, output_tensor):
    logits = output_tensor.contiguous().float()
    loss = F.cross_entropy(logits, labels)
    acc = (logits.argmax(-1) == labels).float().mean()
    rep = torch.stack([loss.detach(), acc])
    g = mpu.get_data_parallel_group()
    if g is not None and not isinstance(g, mpu._SingleGroup) and torch.distributed.get_world_size(g) > 1:
        torch.distributed.all_reduce(rep, group=g)
        rep /= torch.distributed.get_world_size(g)
    return loss, {"loss": rep[0], "accuracy": rep[1]}


def forward_step(data_iterator, model):
    dev = next(model.parameters()).device
    d = next(data_iterator)
    out = model(d["image"].to(dev))
    return out, partial(loss_func, d["label"].to(dev))


if __name__ == "__main__":
    train_valid_test_datasets_provider.is_distributed = True
    pretrain(train_valid_test_datasets_provider, model_provider, "encoder_or_decoder", forward_step,
             extra_args_provider=add_vision_args, args_defaults={"tokenizer_type": "NullTokenizer", "vocab_size": 8})
