"""Legacy vision research models (Swin / MiT backbones, DINO, in-painting)."""
import torch

from megatronapp_b200.legacy.model.vision import DINOHead, DINOLoss, MitInpaintingModel, SwinTransformer, mit_b0
from megatronapp_b200.legacy.model.vision.inpainting import masked_reconstruction_loss
from megatronapp_b200.legacy.model.vision.swin_backbone import SwinBlock, window_partition, window_reverse


def test_window_partition_roundtrip():
    x = torch.randn(2, 14, 14, 5)
    assert torch.equal(window_reverse(window_partition(x, 7), 7, 14, 14), x)


def test_swin_stage_shapes_and_shift_mask_isolates_wrapped_regions():
    net = SwinTransformer(img_size=56, embed_dim=32, depths=(2, 2), num_heads=(2, 4)).eval()
    f = net(torch.randn(2, 3, 56, 56))
    assert [tuple(t.shape[1:]) for t in f] == [(32, 14, 14), (64, 7, 7)]
    blk = SwinBlock(16, (14, 14), 2, ws=7, shift=3).eval()
    # tokens that only meet through the cyclic roll must not attend to each other: perturbing the top-left corner
    # leaves the bottom-right corner (same shifted window, different mask region) unchanged
    x = torch.randn(1, 14 * 14, 16)
    y0 = blk(x)
    x2 = x.clone()
    x2[0, 0, 3] += 10.0
    y1 = blk(x2)
    assert torch.allclose(y0[0, -1], y1[0, -1], atol=1e-6) and not torch.allclose(y0[0, 1], y1[0, 1], atol=1e-4)


def test_mit_pyramid_and_inpainting_output():
    f = mit_b0().eval()(torch.randn(1, 3, 64, 64))
    assert [tuple(t.shape[1:]) for t in f] == [(32, 16, 16), (64, 8, 8), (160, 4, 4), (256, 2, 2)]
    m = MitInpaintingModel().eval()
    img = torch.rand(2, 3, 32, 32)
    mask = torch.zeros(2, 1, 32, 32)
    mask[:, :, 8:16] = 1
    out = m(img * (1 - mask))
    assert out.shape == img.shape
    l1, psnr = masked_reconstruction_loss(out, img, mask)
    assert l1 > 0 and torch.isfinite(psnr)
    assert masked_reconstruction_loss(img, img, mask)[0] == 0


def test_dino_loss_trains_student_towards_teacher_and_tracks_center():
    torch.manual_seed(0)
    head_s, head_t = DINOHead(8, 16, 12, 6), DINOHead(8, 16, 12, 6)
    crit = DINOLoss(16, ncrops=4, warmup_epochs=2, nepochs=4)
    feats = torch.randn(4 * 3, 8)                       # 4 crops x batch 3
    opt = torch.optim.SGD([p for p in head_s.parameters() if p.requires_grad], lr=0.5)
    with torch.no_grad():
        t = head_t(feats[:6])
    losses = []
    for _ in range(20):
        loss = crit(head_s(feats), t, epoch=0)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]
    assert crit.center.abs().sum() > 0
    assert abs(float(crit.teacher_temp_schedule[0]) - 0.04) < 1e-6 and abs(float(crit.teacher_temp_schedule[-1]) - 0.07) < 1e-6
