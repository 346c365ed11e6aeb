"""A scene whose visible splats include large ones centred well OUTSIDE the screen (beyond 1.5x in NDC):
preprocess_fwd does not prefetch their SH rows (its candidate test is "near plane passed and centre within
1.5x the screen") and must fetch them through its fallback once they turn out to touch tiles."""
import numpy as np
import torch

import surfel_scenes as S


def offscreen_scene(P=400, W=128, H=96, seed=31, n_far=120):
    cam = S.make_camera(W, H)
    scene = S.make_scene(P, W, H, seed, depth_complexity=10)
    g = torch.Generator("cpu").manual_seed(seed + 1)
    z = scene["means3D"][:n_far, 2].clamp_min(2.0)
    side = torch.where(torch.rand(n_far, generator=g) < 0.5, -1.0, 1.0)
    horiz = torch.rand(n_far, generator=g) < 0.5
    off = (1.7 + 0.7 * torch.rand(n_far, generator=g)) * side
    inside = (torch.rand(n_far, generator=g) * 2 - 1) * 0.8
    x_ndc = torch.where(horiz, off, inside)
    y_ndc = torch.where(horiz, inside, off)
    scene["means3D"][:n_far, 0] = x_ndc * cam["tanfovx"] * z
    scene["means3D"][:n_far, 1] = y_ndc * cam["tanfovy"] * z
    scene["means3D"][:n_far, 2] = z
    f_pix = W / (2.0 * cam["tanfovx"])
    sigma_px = 40.0 + 60.0 * torch.rand(n_far, 2, generator=g)                 # 3 sigma reaches far into the screen
    scene["scales"][:n_far] = z[:, None] * sigma_px / f_pix
    # mostly camera-facing so that the conic stays bounded (a few random ones are kept)
    q = torch.tensor([1.0, 0.0, 0.0, 0.0]).repeat(n_far, 1) + 0.15 * torch.randn(n_far, 4, generator=g)
    scene["rotations"][:n_far] = torch.nn.functional.normalize(q)
    scene["opacities"][:n_far] = 0.3 + 0.6 * torch.rand(n_far, 1, generator=g)
    return S.to_numpy(scene), S.to_numpy(cam)


def centre_outside_margin(scene, cam, margin=1.5):
    """numpy twin of the kernel's candidate test: True where the projected centre is outside margin x screen."""
    p = np.concatenate([scene["means3D"], np.ones((scene["means3D"].shape[0], 1), np.float32)], 1).astype(np.float32)
    h = p @ cam["projmatrix"].astype(np.float32)
    return (np.abs(h[:, 0]) > margin * np.abs(h[:, 3])) | (np.abs(h[:, 1]) > margin * np.abs(h[:, 3]))
