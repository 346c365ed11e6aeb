"""Kernel lab (run on a GPU box): per-stage device times of the op on a BASELINE workload, for one or more
builds of the library.

  python profiles/kernel_lab.py [--workload headline] [--steps 30] [--libs default,NAME,...]

Each library is measured in its own subprocess (ctypes cannot unload a CUDA library), with CUDA events
recorded around every kernel on the launching stream (surfel_profile_* of the C ABI).  NAME is a variant
built by profiles/build_variants.py into 2d-gaussian-splatting_b200/lib/variants/NAME.so.  The first
library's gradients are saved and every other library is compared with them (max scaled difference), so an
A/B never trades correctness for speed unnoticed.  Prints one JSON line per library."""
import argparse
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "2d-gaussian-splatting_b200")


def child(args):
    sys.path.insert(0, ROOT); sys.path.insert(0, PKG)
    import numpy as np
    import torch
    import surfel_scenes as S
    import diff_surfel_rasterization as dsr
    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer, _cabi
    lib = _cabi.load()
    dev = torch.device("cuda", 0)
    P, W, H = S.CONFIGS[args.workload]
    if args.splats:
        P = args.splats
    scene, cam = S.named(args.workload, P=P)
    gc, go = S.make_cotangents(W, H, S.CONFIG_SEED[args.workload])
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev), sh_degree=3,
        campos=cam["campos"].to(dev), prefiltered=False, debug=False)
    rast = GaussianRasterizer(rs)
    leaf = {k: v.to(dev).requires_grad_(True) for k, v in scene.items()}
    m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
    gc, go = gc.to(dev), go.to(dev)

    def step():
        for t in list(leaf.values()) + [m2d]:
            t.grad = None
        color, radii, allmap = rast(means3D=leaf["means3D"], means2D=m2d, shs=leaf["shs"], opacities=leaf["opacities"],
                                    scales=leaf["scales"], rotations=leaf["rotations"])
        torch.autograd.backward([color, allmap], [gc, go])
        return color, allmap

    for _ in range(args.warmup):
        color, allmap = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / args.steps
    n = lib.surfel_profile_num_stages()
    ms, cnt = (ctypes.c_double * n)(), (ctypes.c_int * n)()
    lib.surfel_profile_enable(1); lib.surfel_profile_read(ms, cnt)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    lib.surfel_profile_enable(0); lib.surfel_profile_read(ms, cnt)
    stages = {lib.surfel_profile_stage_name(i).decode(): round(ms[i] / args.steps, 4) for i in range(n) if cnt[i]}
    out = {"lib": args.tag, "workload": args.workload, "P": P, "ms_per_step": round(ms_step, 4),
           "Msplats_per_s": round(P / ms_step / 1e3, 1), "R": int(dsr.last_num_rendered()), "stage_ms": stages}
    res = {"color": color.detach(), "allmap": allmap.detach(), "means3D": leaf["means3D"].grad, "shs": leaf["shs"].grad,
           "opacities": leaf["opacities"].grad, "scales": leaf["scales"].grad, "rotations": leaf["rotations"].grad}
    ref_path = os.path.join("/tmp", f"lab_ref_{args.workload}_{P}.pt")      # first library's results (not brought back)
    if args.save_ref:
        torch.save({k: v.cpu() for k, v in res.items()}, ref_path)
    elif os.path.exists(ref_path):
        ref = torch.load(ref_path)
        diff = {}
        for k, v in res.items():
            r = ref[k].to(dev)
            diff[k] = float(((v - r).abs() / (r.abs() + 1e-3 * r.abs().max())).max())
        out["max_scaled_diff_vs_first"] = {k: float(f"{d:.2e}") for k, d in diff.items()}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="headline")
    ap.add_argument("--splats", type=int, default=0)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--libs", default="default")
    ap.add_argument("--tag", default=None)
    ap.add_argument("--save-ref", action="store_true")
    args = ap.parse_args()
    if args.tag is not None:
        return child(args)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for i, name in enumerate(args.libs.split(",")):
        env = dict(os.environ)
        if name != "default":
            env["SURFEL_LIB"] = os.path.join(PKG, "lib", "variants", name + ".so")
        cmd = [sys.executable, os.path.abspath(__file__), "--workload", args.workload, "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--tag", name, "--splats", str(args.splats)] + (["--save-ref"] if i == 0 else [])
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        sys.stdout.write(r.stdout if r.returncode == 0 else json.dumps({"lib": name, "error": r.stderr[-800:]}) + "\n")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
