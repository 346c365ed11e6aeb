"""Generates tests/golden/*.npz by IMPORTING THE REFERENCE'S OWN PYTHON (run in the build container,
where /root/reference exists; the fixtures travel, the reference does not).

What the reference can pin (it ships no rasterizer source and no tests — SURVEY §0, §8c):
  ref_conventions.npz
    * T (cov3D_precomp) exactly as /root/reference/gaussian_renderer/__init__.py:64-75 computes it
      when pipe.compute_cov3D_python=True, captured at the rasterizer call (:97-106) by running the
      reference's unmodified render() against a capturing stub of diff_surfel_rasterization;
    * SH -> RGB as :84-91 (eval_sh from /root/reference/utils/sh_utils.py:57-112, +0.5, clamp);
    * camera matrices from /root/reference/scene/cameras.py:50-59 (Camera class);
    * activations feeding the op: /root/reference/scene/gaussian_model.py:95-115.
  oracle_config1.npz
    * regression anchor of the oracle itself on BASELINE config 1 (1k surfels, 256x256): SHA-256 of
      the integer-valued outputs, plus the image outputs for an allclose check.

Usage:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "2d-gaussian-splatting_b200"))
REF = "/root/reference"


def cpu_patches():
    """The reference hard-codes device='cuda'; redirect to CPU for fixture generation."""
    torch.Tensor.cuda = lambda self, *a, **k: self
    for name in ("zeros", "zeros_like", "ones", "ones_like", "tensor", "arange", "empty"):
        orig = getattr(torch, name)

        def wrap(*a, _orig=orig, **k):
            if "device" in k:
                k["device"] = "cpu"
            return _orig(*a, **k)
        setattr(torch, name, wrap)


def stub_modules(capture):
    ply = types.ModuleType("plyfile"); ply.PlyData = object; ply.PlyElement = object
    sys.modules["plyfile"] = ply
    knn = types.ModuleType("simple_knn"); knn_c = types.ModuleType("simple_knn._C"); knn_c.distCUDA2 = None
    sys.modules["simple_knn"] = knn; sys.modules["simple_knn._C"] = knn_c
    for name in ("matplotlib", "matplotlib.pyplot", "cv2"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    dsr = types.ModuleType("diff_surfel_rasterization")

    class Settings:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    class Rasterizer:
        def __init__(self, raster_settings):
            self.rs = raster_settings

        def __call__(self, **kw):
            capture["settings"] = self.rs
            capture["args"] = kw
            H, W = self.rs.image_height, self.rs.image_width
            P = kw["means3D"].shape[0]
            return torch.zeros(3, H, W), torch.zeros(P, dtype=torch.int32), torch.ones(7, H, W)
    dsr.GaussianRasterizationSettings = Settings
    dsr.GaussianRasterizer = Rasterizer
    sys.modules["diff_surfel_rasterization"] = dsr


def main():
    import surfel_scenes as S
    cpu_patches()
    capture = {}
    stub_modules(capture)
    sys.path.insert(0, REF)
    from gaussian_renderer import render
    from scene.cameras import Camera
    from scene.gaussian_model import GaussianModel
    from utils.sh_utils import eval_sh

    W, H, P = 160, 96, 64
    Rm = S.look_at_rotation(20, -10)
    tv = np.array([0.2, -0.1, 0.5])
    mycam = S.make_camera(W, H, R=Rm, t=tv)
    cam = Camera(colmap_id=0, R=Rm, T=tv, FoVx=mycam["FoVx"], FoVy=mycam["FoVy"], image=torch.zeros(3, H, W),
                 gt_alpha_mask=None, image_name="g", uid=0, data_device="cpu")
    scene = S.make_scene(P, W, H, 42, depth_complexity=10)
    m = torch.cat([scene["means3D"], torch.ones(P, 1)], 1) @ mycam["viewmatrix"].inverse()
    scene["means3D"] = m[:, :3].contiguous()

    pc = GaussianModel(3)
    pc.active_sh_degree = 3
    pc._xyz = scene["means3D"]
    pc._scaling = torch.log(scene["scales"])
    pc._rotation = scene["rotations"] * 1.7           # un-normalised on purpose: activation normalises
    pc._opacity = torch.log(scene["opacities"] / (1 - scene["opacities"]))
    pc._features_dc = scene["shs"][:, :1].contiguous()
    pc._features_rest = scene["shs"][:, 1:].contiguous()
    pipe = types.SimpleNamespace(compute_cov3D_python=True, convert_SHs_python=False, depth_ratio=0.0, debug=False)
    render(cam, pc, pipe, torch.zeros(3))
    ref_T = capture["args"]["cov3D_precomp"].detach().numpy().astype(np.float32)
    rs = capture["settings"]

    shs_view = pc.get_features.transpose(1, 2).view(-1, 3, 16)
    dirs = pc.get_xyz - cam.camera_center.repeat(P, 1)
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    ref_rgb = {}
    for deg in range(4):
        ref_rgb[deg] = torch.clamp_min(eval_sh(deg, shs_view, dirs) + 0.5, 0.0).numpy().astype(np.float32)

    np.savez(os.path.join(HERE, "ref_conventions.npz"),
             W=W, H=H, cam_R=Rm, cam_t=tv, fovx=mycam["FoVx"], fovy=mycam["FoVy"],
             means3D=pc.get_xyz.numpy(), scales=pc.get_scaling.numpy(), rotations=pc.get_rotation.numpy(),
             opacities=pc.get_opacity.numpy(), shs=pc.get_features.detach().numpy(),
             ref_T=ref_T, ref_rgb0=ref_rgb[0], ref_rgb1=ref_rgb[1], ref_rgb2=ref_rgb[2], ref_rgb3=ref_rgb[3],
             ref_viewmatrix=cam.world_view_transform.numpy(), ref_projmatrix=cam.full_proj_transform.numpy(),
             ref_campos=cam.camera_center.numpy(), ref_tanfovx=np.float64(rs.tanfovx), ref_tanfovy=np.float64(rs.tanfovy))
    print("wrote ref_conventions.npz: T", ref_T.shape)

    # ---- oracle regression anchor on BASELINE config 1 ----
    from oracle import surfel_oracle as O
    O.build(force=True)
    sc, cm = S.named("config1")
    scn, cmn = S.to_numpy(sc), S.to_numpy(cm)
    bg = np.zeros(3, np.float32)
    pre, binned, img = O.forward(scn, cmn, bg)
    gc, go = S.make_cotangents(cm["W"], cm["H"], S.CONFIG_SEED["config1"])
    grads = O.backward(scn, cmn, bg, pre, binned, img, gc.numpy(), go.numpy())                           # upstream low-pass depth gradient (default)
    grads_x = O.backward(scn, cmn, bg, pre, binned, img, gc.numpy(), go.numpy(), lowpass_quirk=False)     # exact derivative (opt-in)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    np.savez_compressed(
        os.path.join(HERE, "oracle_config1.npz"),
        sha_radii=sha(pre["radii"]), sha_tiles=sha(pre["tiles_touched"]), sha_keys=sha(binned["keys_sorted"]),
        sha_vals=sha(binned["vals_sorted"]), sha_ranges=sha(binned["ranges"]), R=binned["R"],
        color=img["color"].astype(np.float16), others=img["others"].astype(np.float32)[:, ::4, ::4],
        n_contrib_sum=np.int64(img["n_contrib"][0].astype(np.int64).sum()),
        dL_dmeans3D=grads["dL_dmeans3D"], dL_dopacity=grads["dL_dopacity"], dL_dscales=grads["dL_dscales"],
        dL_dmeans2D=grads["dL_dmeans2D"],
        exact_dL_dmeans3D=grads_x["dL_dmeans3D"], exact_dL_dscales=grads_x["dL_dscales"])
    print("wrote oracle_config1.npz: R", binned["R"])


if __name__ == "__main__":
    main()
