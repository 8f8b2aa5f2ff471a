"""DINO self-supervised vision pre-training (reference ``pretrain_vision_dino.py``): multi-crop batches (2 global + N local
views), student / EMA-teacher with the centred cross-entropy; the teacher is updated after every optimizer step by the
training loop (``--vision-pretraining --vision-pretraining-type dino``)."""
import os

if not os.path.exists("/dev/nvidia0"):
    os.environ.setdefault("OMP_NUM_THREADS", "2")

from functools import partial

import torch
import torch.nn.functional as F

from megatronapp_b200.legacy.model.vision.dino import DINOLoss, DINOPretrainModel
from megatronapp_b200.parallel import state as mpu
from megatronapp_b200.training.arguments import core_transformer_config_from_args
from megatronapp_b200.training.global_vars import get_args
from megatronapp_b200.training.training import pretrain
from megatronapp_b200.training.utils import print_rank_0, unwrap_model

_LOSS = None


def add_dino_args(parser):
    g = parser.add_argument_group("dino")
    g.add_argument("--dino-out-dim", type=int, default=65536)
    g.add_argument("--dino-momentum-teacher", type=float, default=0.996)
    g.add_argument("--dino-epochs", type=int, default=100)
    return parser


def model_provider(pre_process=True, post_process=True):
    a = get_args()
    print_rank_0(f"building DINO {a.vision_backbone_type} student / teacher ...")
    cfg = core_transformer_config_from_args(a)
    return DINOPretrainModel(cfg, out_dim=a.dino_out_dim, img_size=a.img_h, local_img_size=a.dino_local_img_size,
                             patch_dim=a.patch_dim, backbone=a.vision_backbone_type, norm_last_layer=a.dino_norm_last_layer,
                             momentum_teacher=a.dino_momentum_teacher, epochs=a.dino_epochs, iters_per_epoch=a.iter_per_epoch,
                             head_hidden=a.dino_head_hidden_size, bottleneck=a.dino_bottleneck_size,
                             swin_type=a.swin_backbone_type)


class MockMultiCropDataset(torch.utils.data.Dataset):
    """Random crops of a procedurally generated image (class dependent colour blobs): 2 global views at ``img`` and
    ``n_local`` local views at ``local`` resolution, each with a random flip and brightness jitter -- the augmentation
    structure of the reference's ``DinoTransform`` without torchvision."""

    def __init__(self, n, img, local, n_local, classes, seed):
        self.n, self.img, self.local, self.n_local, self.classes, self.seed = n, img, local, n_local, classes, seed

    def __len__(self):
        return self.n

    def _view(self, base, size, lo, hi, g):
        H = base.shape[-1]
        s = max(2, int(H * float(torch.empty(1).uniform_(lo, hi, generator=g)) ** 0.5))
        y, x = (int(torch.randint(0, H - s + 1, (1,), generator=g)) for _ in range(2))
        v = F.interpolate(base[None, :, y:y + s, x:x + s], size=(size, size), mode="bilinear", align_corners=False)[0]
        if float(torch.rand(1, generator=g)) < 0.5:
            v = v.flip(-1)
        return v * float(torch.empty(1).uniform_(0.8, 1.2, generator=g))

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        label = int(torch.randint(0, self.classes, (1,), generator=g))
        base = torch.rand(3, 2 * self.img, 2 * self.img, generator=g) * 0.1
        base[label % 3] += 0.5 + 0.4 * (label / max(1, self.classes - 1))
        crops = [self._view(base, self.img, 0.4, 1.0, g) for _ in range(2)]
        crops += [self._view(base, self.local, 0.05, 0.4, g) for _ in range(self.n_local)]
        return {"global": torch.stack(crops[:2]), "local": torch.stack(crops[2:]) if self.n_local else torch.zeros(0),
                "label": torch.tensor(label)}


class _MultiCropView(torch.utils.data.Dataset):
    def __init__(self, ds, n, n_local, local_size):
        self.ds, self.n, self.n_local, self.local = ds, max(int(n), len(ds)), n_local, local_size

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        crops, label = self.ds[i % len(self.ds)]
        if not isinstance(crops, list):                             # evaluation transform: one centre crop, used twice
            crops = [crops, crops]
        return {"global": torch.stack(crops[:2]), "local": torch.stack(crops[2:]) if len(crops) > 2 else torch.zeros(0),
                "label": torch.tensor(label)}


def train_valid_test_datasets_provider(num_samples):
    a = get_args()
    if a.data_path:
        from megatronapp_b200.legacy.data.vit_dataset import build_train_valid_datasets

        tr, _ = build_train_valid_datasets(a.data_path if len(a.data_path) > 1 else a.data_path[0], a.img_h, "dino",
                                           dino_local_img_size=a.dino_local_img_size, dino_local_crops_number=a.dino_local_crops_number,
                                           classes_fraction=a.classes_fraction, data_per_class_fraction=a.data_per_class_fraction)
        mk = lambda n: _MultiCropView(tr, n, a.dino_local_crops_number, a.dino_local_img_size)  # noqa: E731
        return mk(num_samples[0]), mk(num_samples[1]), mk(num_samples[2])
    mk = lambda n, s: MockMultiCropDataset(max(int(n), 1), a.img_h, a.dino_local_img_size, a.dino_local_crops_number,  # noqa: E731
                                           a.num_classes, s)
    return mk(num_samples[0], 1), mk(num_samples[1], 2), mk(num_samples[2], 3)


def loss_func(model, output):
    global _LOSS
    a = get_args()
    student, teacher = output
    if _LOSS is None:
        _LOSS = DINOLoss(student.shape[-1], 2 + a.dino_local_crops_number, a.dino_warmup_teacher_temp, a.dino_teacher_temp,
                         a.dino_warmup_teacher_temp_epochs, a.dino_epochs).to(student.device)
    it = getattr(a, "curr_iteration", 0)
    epoch = it // max(1, a.iter_per_epoch)
    loss = _LOSS(student.float(), teacher.float(), epoch)
    if it < a.dino_freeze_last_layer * a.iter_per_epoch:
        # the weight-normalised last layer does not train during the first epoch(s) (reference cancel_gradients_last_layer)
        for p in unwrap_model(model).student_head.last.parameters():
            p.register_hook(torch.zeros_like) if p.requires_grad and not getattr(p, "_dino_frozen", False) else None
            p._dino_frozen = True
    rep = loss.detach().clone().view(1)
    grp = mpu.get_data_parallel_group()
    if grp is not None and not isinstance(grp, mpu._SingleGroup) and torch.distributed.get_world_size(grp) > 1:
        torch.distributed.all_reduce(rep, group=grp)
        rep /= torch.distributed.get_world_size(grp)
    return loss, {"loss": rep[0]}


def forward_step(data_iterator, model):
    dev = next(model.parameters()).device
    d = next(data_iterator)
    g, loc = d["global"].to(dev), d["local"].to(dev)              # [b, 2, 3, H, W], [b, n, 3, h, w]
    crops = [g[:, 0], g[:, 1]] + [loc[:, i] for i in range(loc.shape[1])] if loc.dim() == 5 else [g[:, 0], g[:, 1]]
    dt = next(model.parameters()).dtype
    return model([c.to(dt) for c in crops]), partial(loss_func, model)


if __name__ == "__main__":
    train_valid_test_datasets_provider.is_distributed = True
    pretrain(train_valid_test_datasets_provider, model_provider, "encoder_or_decoder", forward_step,
             extra_args_provider=add_dino_args,
             args_defaults={"tokenizer_type": "NullTokenizer", "vocab_size": 8, "vision_pretraining": True,
                            "vision_pretraining_type": "dino"})
