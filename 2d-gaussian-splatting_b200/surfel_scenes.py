"""Deterministic synthetic surfel scenes and cameras.

Used by bench.py, __graft_entry__.smoke() and tests/.  Everything is generated on the CPU with a
seeded torch.Generator in float32, exactly as §8(d) prescribes, so the same scene exists in this
container (oracle / golden fixtures) and on the GPU box.

Camera conventions restate the reference's own helpers (checked against them in
tests/test_golden.py when /root/reference is present):
  getWorld2View2 / getProjectionMatrix  /root/reference/utils/graphics_utils.py:38-71
  world_view_transform / full_proj_transform / camera_center  /root/reference/scene/cameras.py:56-59
"""
import math

import numpy as np
import torch

# name -> (P, W, H); BASELINE.json configs + the headline metric config
CONFIGS = {
    "config1": (1_000, 256, 256),
    "config2": (100_000, 1920, 1080),
    "config3": (1_000_000, 1600, 1200),
    "headline": (1_000_000, 1920, 1080),
    "config4": (5_000_000, 3840, 2160),
    "config5": (2_000_000, 7680, 4320),
}
CONFIG_SEED = {"config1": 1, "config2": 2, "config3": 3, "headline": 0, "config4": 4, "config5": 5}


def projection_matrix(znear, zfar, fovX, fovY):
    """OpenGL-style perspective with z_sign=+1 (reference utils/graphics_utils.py:51-71)."""
    tanHalfFovY, tanHalfFovX = math.tan(fovY / 2), math.tan(fovX / 2)
    top, right = tanHalfFovY * znear, tanHalfFovX * znear
    bottom, left = -top, -right
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def world2view(R, t, translate=np.array([0.0, 0.0, 0.0]), scale=1.0):
    """W2C with R stored transposed, COLMAP convention, including the reference's inverse /
    re-inverse round trip so the float32 result is bit-identical
    (reference utils/graphics_utils.py:38-49, getWorld2View2)."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = np.linalg.inv(Rt)
    cam_center = (C2W[:3, 3] + translate) * scale
    C2W[:3, 3] = cam_center
    Rt = np.linalg.inv(C2W)
    return np.float32(Rt)


def make_camera(W, H, fovy_deg=50.0, R=None, t=None, znear=0.01, zfar=100.0):
    """Returns the dict of per-view inputs in the layout scene/cameras.py:56-59 produces."""
    R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64)
    t = np.zeros(3) if t is None else np.asarray(t, dtype=np.float64)
    fovy = math.radians(fovy_deg)
    tanfovy = math.tan(fovy / 2)
    tanfovx = tanfovy * W / H
    fovx = 2 * math.atan(tanfovx)
    wvt = torch.tensor(world2view(R, t)).transpose(0, 1).contiguous()
    proj = projection_matrix(znear, zfar, fovx, fovy).transpose(0, 1)
    full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0).contiguous()
    campos = wvt.inverse()[3, :3].contiguous()
    return dict(W=W, H=H, tanfovx=tanfovx, tanfovy=tanfovy, FoVx=fovx, FoVy=fovy,
                viewmatrix=wvt, projmatrix=full, campos=campos, znear=znear, zfar=zfar)


def look_at_rotation(yaw_deg, pitch_deg):
    cy, sy = math.cos(math.radians(yaw_deg)), math.sin(math.radians(yaw_deg))
    cp, sp = math.cos(math.radians(pitch_deg)), math.sin(math.radians(pitch_deg))
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    return Ry @ Rx


def make_scene(P, W, H, seed, fovy_deg=50.0, depth_complexity=30.0, sh_coeffs=16, sigma_scale=1.0):
    """§8(d) generator.  Returns CPU float32 tensors: means3D (P,3), scales (P,2), rotations (P,4),
    opacities (P,1), shs (P,16,3).  Positions are in the frame of the identity camera.
    Generated on ONE thread: PyTorch's vectorised CPU kernels (exp, sigmoid, norm) round the elements at the
    seams of its per-thread chunks differently, so with the intra-op thread count of the moment the "same"
    scene differed in the last bit of a few values — enough to flip alpha >= 1/255 decisions between two
    processes (seen as a 1e-2-sized difference between two A/B runs in profiles/kernel_lab.py)."""
    nthreads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        return _make_scene(P, W, H, seed, fovy_deg, depth_complexity, sh_coeffs, sigma_scale)
    finally:
        torch.set_num_threads(nthreads)


def _make_scene(P, W, H, seed, fovy_deg, depth_complexity, sh_coeffs, sigma_scale):
    g = torch.Generator("cpu").manual_seed(1234 + seed)

    def U(n, lo, hi):
        return torch.rand(n, generator=g) * (hi - lo) + lo

    tanfovy = math.tan(math.radians(fovy_deg) / 2)
    tanfovx = tanfovy * W / H
    z = U(P, 2.0, 12.0)
    behind = torch.rand(P, generator=g) < 0.02
    z = torch.where(behind, U(P, -1.0, 0.2), z)
    x = U(P, -1.05, 1.05) * tanfovx * z
    y = U(P, -1.05, 1.05) * tanfovy * z
    means3D = torch.stack([x, y, z], 1).contiguous()
    f_pix = W / (2 * tanfovx)
    sigma_med = sigma_scale * math.sqrt(depth_complexity * W * H / (9 * math.pi * P * math.exp(0.25)))
    sigma = torch.exp(torch.randn(P, 2, generator=g) * 0.5 + math.log(sigma_med))
    scales = (z.abs().clamp_min(0.2)[:, None] * sigma / f_pix).contiguous()
    q = torch.randn(P, 4, generator=g)
    rotations = (q / q.norm(dim=1, keepdim=True)).contiguous()
    opacities = torch.sigmoid(torch.randn(P, 1, generator=g) * 2.0).contiguous()
    shs = torch.cat([torch.randn(P, 1, 3, generator=g),
                     torch.randn(P, sh_coeffs - 1, 3, generator=g) * 0.2], 1).contiguous()
    return dict(means3D=means3D, scales=scales, rotations=rotations, opacities=opacities, shs=shs)


def make_cotangents(W, H, seed):
    g = torch.Generator("cpu").manual_seed(1234 + seed + 100)
    return torch.randn(3, H, W, generator=g), torch.randn(7, H, W, generator=g)


def named(name, P=None):
    """Scene + camera of a BASELINE.json config (optionally with P overridden for small tests)."""
    P0, W, H = CONFIGS[name]
    P = P0 if P is None else P
    cam = make_camera(W, H)
    scene = make_scene(P, W, H, CONFIG_SEED[name])
    return scene, cam


def to_numpy(d):
    return {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in d.items()}
