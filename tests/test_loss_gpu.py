"""Fused L1+SSIM loss against a PyTorch restatement of the reference's
l1_loss / ssim (/root/reference/utils/loss_utils.py:6-7, :43-73; combined at train.py:73-74)."""
from math import exp

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def reference_loss(img, gt, lam):
    g = torch.tensor([exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    window = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0).expand(3, 1, 11, 11).contiguous().to(img.device)
    conv = lambda t: F.conv2d(t, window, padding=5, groups=3)
    mu1, mu2 = conv(img), conv(gt)
    s1, s2, s12 = conv(img * img) - mu1.pow(2), conv(gt * gt) - mu2.pow(2), conv(img * gt) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim = (((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1.pow(2) + mu2.pow(2) + C1) * (s1 + s2 + C2))).mean()
    return (1.0 - lam) * torch.abs(img - gt).mean() + lam * (1.0 - ssim)


@pytest.mark.parametrize("shape,lam", [((3, 97, 131), 0.2), ((3, 256, 320), 0.2), ((3, 64, 48), 1.0), ((3, 33, 17), 0.0)])
def test_fused_l1_ssim_matches_reference(cuda_lib, shape, lam):
    from diff_surfel_rasterization.loss import l1_ssim_loss
    g = torch.Generator("cpu").manual_seed(shape[1])
    base = torch.rand(*shape, generator=g)
    gt = (base + 0.1 * torch.randn(*shape, generator=g)).clamp(0, 1).cuda()
    img0 = (base + 0.15 * torch.randn(*shape, generator=g)).clamp(0, 1).cuda()
    res = {}
    for name, fn in (("ref", reference_loss), ("fused", l1_ssim_loss)):
        img = img0.clone().requires_grad_(True)
        loss = fn(img, gt, lam)
        (loss * 3.0).backward()
        res[name] = (float(loss.detach()), img.grad)
    assert abs(res["ref"][0] - res["fused"][0]) < 2e-5 * max(1.0, abs(res["ref"][0])), (res["ref"][0], res["fused"][0])
    gr, gf = res["ref"][1], res["fused"][1]
    err = (gr - gf).abs().max() / gr.abs().max().clamp_min(1e-12)
    assert float(err) < 2e-4, float(err)
