"""Timeline of the host-buffer pipeline (surfel_host.HostStepPipeline): per step, when its H2D, compute
and D2H start and end (CUDA events on the three streams), to see which stream bounds the e2e leg.

  python profiles/e2e_timeline.py [--steps 12]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "2d-gaussian-splatting_b200"))
import torch

import bench
import surfel_scenes as S
from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from surfel_host import HostStepPipeline, NAMES


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=12)
    args = ap.parse_args()
    bench.bind_to_gpu_numa_node(0)
    dev = torch.device("cuda", 0)
    P, W, H = S.CONFIGS["headline"]
    scene, cam = S.named("headline")
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev), sh_degree=3,
        campos=cam["campos"].to(dev), prefiltered=False, debug=False)
    rast = GaussianRasterizer(rs)
    gc, go = S.make_cotangents(W, H, 5)
    host_in = {k: scene[k].contiguous().pin_memory() for k in NAMES}
    host_gc, host_go = gc.pin_memory(), go.pin_memory()
    host_out = {"color": torch.empty(3, H, W).pin_memory(), "allmap": torch.empty(7, H, W).pin_memory(),
                "radii": torch.empty(P, dtype=torch.int32).pin_memory()}
    host_grad = {k: torch.empty_like(scene[k]).pin_memory() for k in NAMES}
    host_grad["means2D"] = torch.empty(P, 3).pin_memory()
    pipe = HostStepPipeline(rast, host_in, host_gc, host_go, dev)
    pipe.run(3, host_in, host_gc, host_go, host_out, host_grad)
    torch.cuda.synchronize()
    pipe.trace = []
    t0 = torch.cuda.Event(enable_timing=True)
    t0.record(pipe.s_in)
    pipe.run(args.steps, host_in, host_gc, host_go, host_out, host_grad)
    torch.cuda.synchronize()
    rows = {}
    for kind, step, a, b in pipe.trace:
        rows.setdefault(step, {})[kind] = (round(t0.elapsed_time(a), 2), round(t0.elapsed_time(b), 2))
    for step in sorted(rows):
        print(step, json.dumps(rows[step]))
    last = max(v["d2h"][1] for v in rows.values())
    print(json.dumps({"steps": args.steps, "ms_per_step": last / args.steps}))


if __name__ == "__main__":
    main()
