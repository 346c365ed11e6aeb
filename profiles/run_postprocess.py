"""Times the fused post-process (row f1) against the reference's eager PyTorch tail on one GPU at
1920x1080, forward + backward; prints one JSON line.  Usage: python profiles/run_postprocess.py"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2d-gaussian-splatting_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

import surfel_scenes as S
from diff_surfel_rasterization import _cabi
from diff_surfel_rasterization.postprocess import surface_outputs
from test_postprocess_gpu import reference_tail

dev = "cuda"
W, H = 1920, 1080
cam = S.make_camera(W, H)
view = types.SimpleNamespace(world_view_transform=cam["viewmatrix"].to(dev), full_proj_transform=cam["projmatrix"].to(dev),
                             image_width=W, image_height=H)
g = torch.Generator("cpu").manual_seed(0)
allmap0 = torch.rand(7, H, W, generator=g).to(dev) + 0.1
allmap0[0] *= 5; allmap0[5] *= 5
cot = {k: torch.randn(*s, generator=g).to(dev) for k, s in
       dict(rend_alpha=(1, H, W), rend_normal=(3, H, W), rend_dist=(1, H, W), surf_depth=(1, H, W), surf_normal=(3, H, W)).items()}
out = {}
for name, fn in (("eager_pytorch", reference_tail), ("fused_cuda", surface_outputs)):
    def step():
        a = allmap0.clone().requires_grad_(True)
        o = fn(a, view, 1.0)
        torch.autograd.backward([o[k] for k in cot], [cot[k] for k in cot])
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
out["what"] = "render() post-process fwd+bwd at 1920x1080 (includes the allmap clone both ways)"
print(json.dumps(out))
