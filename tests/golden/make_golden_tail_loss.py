"""Pins the PyTorch restatements used by the f1 / f2 parity tests to THE REFERENCE'S OWN PYTHON.

Run in the build container (where /root/reference exists); writes tests/golden/ref_tail_loss.npz.

  * tail: the reference's unmodified gaussian_renderer.render() is run (on CPU) against a stub
    rasterizer that returns a prepared `allmap`; what comes back — rend_alpha, rend_normal, rend_dist,
    surf_depth, surf_normal (/root/reference/gaussian_renderer/__init__.py:118-156,
    /root/reference/utils/point_utils.py:9-37) — and the gradient of a fixed scalar of those maps with
    respect to `allmap` are stored for depth_ratio 0, 1 and 0.3.
  * loss: /root/reference/utils/loss_utils.py l1_loss, ssim and train.py:73-74's combination, value and
    gradient with respect to the image.

Usage:  python tests/golden/make_golden_tail_loss.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "2d-gaussian-splatting_b200"))
REF = "/root/reference"


def make_allmap(W, H, seed):
    """A plausible rasterizer output: alpha in [0,1] with holes, alpha-weighted depth and normals."""
    g = torch.Generator("cpu").manual_seed(seed)
    alpha = torch.rand(1, H, W, generator=g).clamp(0.0, 1.0)
    alpha[torch.rand(1, H, W, generator=g) < 0.12] = 0.0                        # holes: D/alpha = nan -> nan_to_num
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    z = (3.0 + 1.5 * torch.sin(5 * xx) * torch.cos(4 * yy) + 0.3 * torch.rand(H, W, generator=g))[None]
    n = torch.nn.functional.normalize(torch.randn(3, H, W, generator=g), dim=0)
    median = torch.where(alpha > 0.5, z + 0.05 * torch.randn(1, H, W, generator=g), torch.zeros(1, H, W))
    dist = 0.01 * torch.rand(1, H, W, generator=g)
    return torch.cat([alpha * z, alpha, n * alpha, median, dist], 0).contiguous()


def main():
    import make_golden as MG
    import surfel_scenes as S
    MG.cpu_patches()
    holder = {}
    MG.stub_modules({})                                                         # plyfile / simple_knn / cv2 stubs
    dsr = types.ModuleType("diff_surfel_rasterization")

    class Settings:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    class Rasterizer:
        def __init__(self, raster_settings):
            self.rs = raster_settings

        def __call__(self, **kw):
            P = kw["means3D"].shape[0]
            H, W = self.rs.image_height, self.rs.image_width
            return torch.zeros(3, H, W), torch.ones(P, dtype=torch.int32), holder["allmap"]
    dsr.GaussianRasterizationSettings, dsr.GaussianRasterizer = Settings, Rasterizer
    sys.modules["diff_surfel_rasterization"] = dsr
    sys.path.insert(0, REF)
    from gaussian_renderer import render
    from scene.cameras import Camera
    from scene.gaussian_model import GaussianModel
    from utils.loss_utils import l1_loss, ssim

    W, H, P = 48, 36, 8
    Rm = S.look_at_rotation(12, -7)
    tv = np.array([0.15, -0.05, 0.4])
    mycam = S.make_camera(W, H, R=Rm, t=tv)
    cam = Camera(colmap_id=0, R=Rm, T=tv, FoVx=mycam["FoVx"], FoVy=mycam["FoVy"], image=torch.zeros(3, H, W),
                 gt_alpha_mask=None, image_name="g", uid=0, data_device="cpu")
    scene = S.make_scene(P, W, H, 3, depth_complexity=2)
    pc = GaussianModel(3)
    pc.active_sh_degree = 3
    pc._xyz, pc._scaling, pc._rotation = scene["means3D"], torch.log(scene["scales"]), scene["rotations"]
    pc._opacity = torch.log(scene["opacities"] / (1 - scene["opacities"]))
    pc._features_dc, pc._features_rest = scene["shs"][:, :1].contiguous(), scene["shs"][:, 1:].contiguous()

    out = {"W": W, "H": H, "viewmatrix": cam.world_view_transform.numpy(), "projmatrix": cam.full_proj_transform.numpy()}
    allmap0 = make_allmap(W, H, 17)
    out["allmap"] = allmap0.numpy()
    keys = ("rend_alpha", "rend_normal", "rend_dist", "surf_depth", "surf_normal")
    g = torch.Generator("cpu").manual_seed(23)
    cot = {k: torch.randn((3 if "normal" in k else 1), H, W, generator=g) for k in keys}
    for k, v in cot.items():
        out["cot_" + k] = v.numpy()
    for ratio in (0.0, 1.0, 0.3):
        holder["allmap"] = allmap0.clone().requires_grad_(True)
        pipe = types.SimpleNamespace(compute_cov3D_python=False, convert_SHs_python=False, depth_ratio=ratio, debug=False)
        rets = render(cam, pc, pipe, torch.zeros(3))
        scalar = sum((rets[k] * cot[k]).sum() for k in keys)
        scalar.backward()
        tag = str(ratio).replace(".", "p")
        for k in keys:
            out[f"tail_{tag}_{k}"] = rets[k].detach().numpy()
        out[f"tail_{tag}_grad_allmap"] = holder["allmap"].grad.numpy()

    shape = (3, 36, 48)
    base = torch.rand(*shape, generator=g)
    gt = (base + 0.1 * torch.randn(*shape, generator=g)).clamp(0, 1)
    img0 = (base + 0.15 * torch.randn(*shape, generator=g)).clamp(0, 1)
    out["loss_img"], out["loss_gt"] = img0.numpy(), gt.numpy()
    for lam in (0.2, 1.0, 0.0):
        img = img0.clone().requires_grad_(True)
        Ll1 = l1_loss(img, gt)
        loss = (1.0 - lam) * Ll1 + lam * (1.0 - ssim(img, gt))                  # train.py:73-74
        loss.backward()
        tag = str(lam).replace(".", "p")
        out[f"loss_{tag}_value"] = np.float64(loss.item())
        out[f"loss_{tag}_grad"] = img.grad.numpy()
    out["l1_value"] = np.float64(l1_loss(img0, gt).item())
    out["ssim_value"] = np.float64(ssim(img0, gt).item())
    np.savez_compressed(os.path.join(HERE, "ref_tail_loss.npz"), **out)
    print("wrote ref_tail_loss.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
