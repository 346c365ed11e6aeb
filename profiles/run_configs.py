"""Times fwd+bwd of the op on every BASELINE.json config that fits one GPU (and config 5 as one of 8
tile bands), with CUDA events; writes profiles/<tag>_configs.json.  Usage: python profiles/run_configs.py r1"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "2d-gaussian-splatting_b200"))
import torch

import diff_surfel_rasterization as dsr
import surfel_parallel as SP
import surfel_scenes as S
from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer


def run(name, band=None, steps=10, warm=3):
    P, W, H = S.CONFIGS[name]
    dev = torch.device("cuda")
    scene, cam = S.named(name)
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev), sh_degree=3,
        campos=cam["campos"].to(dev), prefiltered=False, debug=False, tile_rows=band)
    leaf = {k: v.to(dev).requires_grad_(True) for k, v in scene.items()}
    m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
    gc, go = S.make_cotangents(W, H, 1)
    gc, go = gc.to(dev), go.to(dev)
    rast = GaussianRasterizer(rs)

    def step():
        for t in list(leaf.values()) + [m2d]:
            t.grad = None
        color, radii, allmap = rast(means3D=leaf["means3D"], means2D=m2d, shs=leaf["shs"], opacities=leaf["opacities"],
                                    scales=leaf["scales"], rotations=leaf["rotations"])
        torch.autograd.backward([color, allmap], [gc, go])
        return color, radii, allmap
    for _ in range(warm):
        color, radii, allmap = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ok = bool(torch.isfinite(color).all() and torch.isfinite(allmap).all() and all(torch.isfinite(t.grad).all() for t in leaf.values()))
    return dict(config=name, P=P, W=W, H=H, band=band, visible=int((radii > 0).sum()), instances=int(dsr.last_num_rendered()),
                ms_fwd_bwd=ms, Msplats_per_s=P / ms / 1e3, finite=ok, peak_mem_GB=torch.cuda.max_memory_allocated() / 1e9)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "rX"
    out = []
    for name in ("config1", "config2", "config3", "headline", "config4"):
        out.append(run(name)); print(out[-1], flush=True)
        torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
    H = S.CONFIGS["config5"][2]
    out.append(run("config5", steps=3, warm=1)); print(out[-1], flush=True)                       # whole 8K frame on one GPU
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
    out.append(run("config5", band=SP.tile_row_band(H, 3, 8), steps=5, warm=2)); print(out[-1], flush=True)  # rank 3 of 8
    json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_configs.json"), "w"), indent=1)
