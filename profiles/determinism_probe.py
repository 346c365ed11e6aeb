"""Stress test for rare races: the forward of the op is bit-deterministic by construction (sorted lists, no
atomics on its outputs), so N forward+backward steps on the same inputs must reproduce the first step's
radii / color / allmap / image state / point list / ranges / records bit for bit.  Any mismatch is localised to
the first stage whose output differs.   python profiles/determinism_probe.py [--iters 300] [--workload headline]"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "2d-gaussian-splatting_b200"))
import torch

import surfel_scenes as S
import diff_surfel_rasterization as dsr
from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer, _RasterizeGaussians, _cabi


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--workload", default="headline")
    ap.add_argument("--splats", type=int, default=0)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    P, W, H = S.CONFIGS[args.workload]
    P = args.splats or P
    scene, cam = S.named(args.workload, P=P)
    gc, go = S.make_cotangents(W, H, 3)
    gc, go = gc.to(dev), go.to(dev)
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev), sh_degree=3,
        campos=cam["campos"].to(dev), prefiltered=False, debug=False)
    leaf = {k: v.to(dev).requires_grad_(True) for k, v in scene.items()}
    m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
    lib = _cabi.load()
    saved = {}

    # capture the workspaces the autograd node saves
    orig = _RasterizeGaussians.backward

    def spy(ctx, *a):
        saved["ws"] = ctx.saved_tensors
        saved["cap"] = ctx.num_rendered
        return orig(ctx, *a)
    _RasterizeGaussians.backward = staticmethod(spy)

    def snapshot():
        for t in list(leaf.values()) + [m2d]:
            t.grad = None
        color, radii, allmap = GaussianRasterizer(rs)(means3D=leaf["means3D"], means2D=m2d, shs=leaf["shs"], opacities=leaf["opacities"],
                                                      scales=leaf["scales"], rotations=leaf["rotations"])
        torch.autograd.backward([color, allmap], [gc, go])
        torch.cuda.synchronize()
        ws = saved["ws"]
        radii_, geom, binning, img = ws[5], ws[6], ws[7], ws[8]
        R = dsr.last_num_rendered()
        go_ = (ctypes.c_size_t * 6)(); lib.surfel_geom_offsets(P, go_)
        bo = (ctypes.c_size_t * 5)(); lib.surfel_binning_offsets(saved["cap"], W, H, bo)
        io = (ctypes.c_size_t * 2)(); lib.surfel_image_offsets(W, H, io)
        vis = radii_ > 0
        rec = geom[go_[0]:go_[0] + P * 128].view(torch.int32).view(P, 32)[vis]
        tm = geom[go_[5]:go_[5] + P * 48].view(torch.int32).view(P, 12)[vis]
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        return {"R": torch.tensor([R]), "radii": radii_.clone(), "records": rec.clone(), "transform": tm.clone(),
                "offsets": geom[go_[2]:go_[2] + 4 * P].clone(),
                "ranges": binning[bo[4]:bo[4] + 8 * tiles].clone(), "point_list": binning[bo[3]:bo[3] + 4 * R].clone(),
                "image_state": img[io[0]:io[0] + 20 * W * H].clone(), "color": color.detach().clone(), "allmap": allmap.detach().clone()}

    ref = snapshot()
    bad = []
    for it in range(args.iters):
        cur = snapshot()
        for k in ref:
            if not torch.equal(ref[k], cur[k]):
                n = int((ref[k] != cur[k]).sum()) if ref[k].shape == cur[k].shape else -1
                bad.append({"iter": it, "first_stage_that_differs": k, "elements": n})
                break
        if len(bad) >= 5:
            break
    print(json.dumps({"workload": args.workload, "P": P, "iters": args.iters, "mismatches": bad,
                      "bit_identical": not bad}))


if __name__ == "__main__":
    main()
