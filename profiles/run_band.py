"""BASELINE config 5: ONE 7680x4320 frame of 2 M surfels, tile-row bands over N GPUs, one NCCL
all-gather of the band outputs.  Launch with torchrun; rank 0 prints one JSON line and
(if --check) verifies the stitched frame against a single-GPU render of the whole frame.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29533 profiles/run_band.py --steps 10 --check
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "2d-gaussian-splatting_b200"))
import torch
import torch.distributed as dist

import surfel_parallel as SP
import surfel_scenes as S
from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="config5")
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    P, W, H = S.CONFIGS[args.workload]
    scene, cam = S.named(args.workload)                       # same seed on every rank: replicated splats
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev), sh_degree=3,
        campos=cam["campos"].to(dev), prefiltered=False, debug=False)
    leaf = {k: v.to(dev).requires_grad_(True) for k, v in scene.items()}
    m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
    gc, go = S.make_cotangents(W, H, 5)
    gc, go = gc.to(dev), go.to(dev)
    names = ("means3D", "scales", "rotations", "opacities", "shs")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    acc = [0.0, 0.0, 0.0, 0.0]

    def step(timed):
        for t in list(leaf.values()) + [m2d]:
            t.grad = None
        ev[0].record()
        band = SP.tile_row_band(H, rank, world)
        color, radii, allmap = GaussianRasterizer(rs._replace(tile_rows=band))(
            means3D=leaf["means3D"], means2D=m2d, shs=leaf["shs"], opacities=leaf["opacities"],
            scales=leaf["scales"], rotations=leaf["rotations"])
        ev[1].record()
        full = SP._BandGather.apply(torch.cat([color, allmap], 0), H, rank, world, None)     # NCCL all-gather
        ev[2].record()
        torch.autograd.backward([full], [torch.cat([gc, go], 0)])                             # slice + band backward
        ev[3].record()
        SP.allreduce_gradients([leaf[k].grad for k in names])                                 # sum of band partials
        ev[4].record()
        torch.cuda.synchronize()
        if timed:
            for i in range(4):
                acc[i] += ev[i].elapsed_time(ev[i + 1])
        return full

    for _ in range(args.warmup):
        full = step(False)
    dist.barrier(); torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        full = step(True)
    t1.record()
    torch.cuda.synchronize()
    tt = torch.tensor([t0.elapsed_time(t1) / args.steps] + [a / args.steps for a in acc], device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    out = {"workload": f"{args.workload}: {P} surfels, {W}x{H}, tile-row bands over {world} GPUs", "n_gpus": world,
           "ms_per_frame_fwd_bwd": float(tt[0]), "ms_band_forward": float(tt[1]), "ms_allgather_outputs": float(tt[2]),
           "ms_band_backward": float(tt[3]), "ms_allreduce_grads": float(tt[4]),
           "gather_bytes_per_rank": int(10 * 4 * W * (H // world)), "Msplats_per_s": P / float(tt[0]) / 1e3}
    if args.check and rank == 0:
        ref_leaf = {k: v.to(dev).requires_grad_(True) for k, v in scene.items()}
        m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        color, radii, allmap = GaussianRasterizer(rs)(means3D=ref_leaf["means3D"], means2D=m2, shs=ref_leaf["shs"],
                                                      opacities=ref_leaf["opacities"], scales=ref_leaf["scales"], rotations=ref_leaf["rotations"])
        torch.autograd.backward([color, allmap], [gc, go])
        e1.record(); torch.cuda.synchronize()
        out["single_gpu_ms_fwd_bwd_cold"] = e0.elapsed_time(e1)
        out["stitched_equals_single_gpu"] = bool(torch.equal(full[:3], color) and torch.equal(full[3:], allmap))
        errs = {k: float((leaf[k].grad - ref_leaf[k].grad).abs().max() / (ref_leaf[k].grad.abs().max() + 1e-30)) for k in names}
        out["max_rel_grad_diff_vs_single_gpu"] = max(errs.values())
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
