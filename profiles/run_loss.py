"""Times the fused L1+SSIM loss (row f2) against the reference's eager PyTorch version at 1920x1080,
forward + backward, on one GPU; prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2d-gaussian-splatting_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

from diff_surfel_rasterization.loss import l1_ssim_loss
from test_loss_gpu import reference_loss

g = torch.Generator("cpu").manual_seed(0)
gt = torch.rand(3, 1080, 1920, generator=g).cuda()
img0 = (gt + 0.1 * torch.randn(3, 1080, 1920, generator=g).cuda()).clamp(0, 1)
out = {}
for name, fn in (("eager_pytorch", reference_loss), ("fused_cuda", l1_ssim_loss)):
    def step():
        img = img0.clone().requires_grad_(True)
        fn(img, gt, 0.2).backward()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        step()
    e1.record(); torch.cuda.synchronize()
    out[name + "_ms"] = e0.elapsed_time(e1) / 30
out["speedup"] = out["eager_pytorch_ms"] / out["fused_cuda_ms"]
out["what"] = "(1-l)*L1 + l*(1-SSIM) fwd+bwd at 3x1080x1920 (includes the image clone both ways)"
print(json.dumps(out))
