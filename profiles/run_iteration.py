"""BASELINE config 3: 1 M surfels, 1600x1200 (DTU shape), depth_ratio = 1, normal + distortion
regularisers, one training-iteration analogue (rasterize -> render() tail -> train.py loss -> backward)
on one B200.  Two variants: the reference's eager PyTorch tail + loss (restated in tests/) on top of
the CUDA op, and the fused tail (f1) + fused loss (f2).  Prints one JSON line."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2d-gaussian-splatting_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

import surfel_scenes as S
from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from diff_surfel_rasterization.loss import l1_ssim_loss
from diff_surfel_rasterization.optim import FusedAdam, densification_stats
from diff_surfel_rasterization.postprocess import surface_outputs
from test_loss_gpu import reference_loss
from test_postprocess_gpu import reference_tail

name = sys.argv[1] if len(sys.argv) > 1 else "config3"
dev = "cuda"
P, W, H = S.CONFIGS[name]
scene, cam = S.named(name)
rs = GaussianRasterizationSettings(
    image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=torch.zeros(3, device=dev),
    scale_modifier=1.0, viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev), sh_degree=3,
    campos=cam["campos"].to(dev), prefiltered=False, debug=False)
view = types.SimpleNamespace(world_view_transform=cam["viewmatrix"].to(dev), full_proj_transform=cam["projmatrix"].to(dev),
                             image_width=W, image_height=H)
leaf = {k: v.to(dev).requires_grad_(True) for k, v in scene.items()}
m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
gt = torch.rand(3, H, W, generator=torch.Generator("cpu").manual_seed(7)).to(dev)
lam_ssim, lam_n, lam_d, ratio = 0.2, 0.05, 1000.0, 1.0
rast = GaussianRasterizer(rs)


accum, denom, maxr = torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, device=dev)
groups = lambda: [{"params": [leaf[k]], "lr": 0.0, "name": k} for k in leaf]   # lr 0: same scene every iteration
optim = {"eager": torch.optim.Adam(groups(), lr=0.0, eps=1e-15), "fused": FusedAdam(groups(), lr=0.0, eps=1e-15)}


def eager_stats(radii):
    vis = radii > 0                                            # train.py:125-128, gaussian_model.py:405-407
    maxr[vis] = torch.max(maxr[vis], radii[vis])
    accum[vis] += torch.norm(m2d.grad[vis], dim=-1, keepdim=True)
    denom[vis] += 1


def iteration(tail, loss_fn, update=None):
    for t in list(leaf.values()) + [m2d]:
        t.grad = None
    image, radii, allmap = rast(means3D=leaf["means3D"], means2D=m2d, shs=leaf["shs"], opacities=leaf["opacities"],
                                scales=leaf["scales"], rotations=leaf["rotations"])
    o = tail(allmap, view, ratio)
    loss = loss_fn(image, gt, lam_ssim)
    normal_loss = lam_n * (1 - (o["rend_normal"] * o["surf_normal"]).sum(dim=0))[None].mean()
    dist_loss = lam_d * o["rend_dist"].mean()
    total = loss + dist_loss + normal_loss
    total.backward()
    if update == "eager":
        eager_stats(radii)
        optim["eager"].step()
    elif update == "fused":
        densification_stats(accum, denom, maxr, m2d.grad, radii)
        optim["fused"].step()
    return total


out = {"workload": f"{name}: {P} surfels, {W}x{H}, depth_ratio=1, L1+DSSIM + normal + distortion, fwd+bwd"}
vals = {}
for tag, tail, lf in (("eager_tail_and_loss", reference_tail, reference_loss), ("fused_tail_and_loss", surface_outputs, l1_ssim_loss)):
    for _ in range(3):
        tot = iteration(tail, lf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        tot = iteration(tail, lf)
    e1.record(); torch.cuda.synchronize()
    out[tag + "_ms_per_iteration"] = e0.elapsed_time(e1) / 20
    vals[tag] = (float(tot.detach()), leaf["means3D"].grad.clone())
for tag, tail, lf, upd in (("eager_with_update", reference_tail, reference_loss, "eager"),
                           ("fused_with_update", surface_outputs, l1_ssim_loss, "fused")):
    for _ in range(3):
        iteration(tail, lf, upd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        iteration(tail, lf, upd)
    e1.record(); torch.cuda.synchronize()
    out[tag + "_ms_per_iteration"] = e0.elapsed_time(e1) / 20
out["iterations_per_s_fused_with_update"] = 1e3 / out["fused_with_update_ms_per_iteration"]
out["loss_eager"], out["loss_fused"] = vals["eager_tail_and_loss"][0], vals["fused_tail_and_loss"][0]
ga, gb = vals["eager_tail_and_loss"][1], vals["fused_tail_and_loss"][1]
out["max_rel_grad_diff_means3D"] = float((ga - gb).abs().max() / ga.abs().max())
out["iterations_per_s_fused"] = 1e3 / out["fused_tail_and_loss_ms_per_iteration"]
print(json.dumps(out))
