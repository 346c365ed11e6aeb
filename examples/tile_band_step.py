"""One oversized frame rendered by N GPUs in tile-row bands, with the exchange of the band outputs done by the
render kernel's own stores (surfel_parallel.rasterize_tile_band, gather="fused") — the multi-GPU mode of
DESIGN.md section 6 as a caller would use it.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/tile_band_step.py

Every rank holds the full splat set; rank r renders band r of the frame straight into every GPU's copy of it
(symmetric memory over NVLink), the per-pixel loss of this example is local to the band, and the gradients come
back summed over the ranks (16 floats per splat are all-reduced; the SH gradient is expanded from the summed
colour gradient afterwards).
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-gaussian-splatting_b200")]


def main():
    import surfel_parallel as SP
    import surfel_scenes as S
    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", torch.cuda.current_device())
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    P, W, H = 2_000_000, 7680, 4320                      # BASELINE config 5
    cam = S.make_camera(W, H)
    scene = S.make_scene(P, W, H, seed=5)                # same seed on every rank: replicated splats
    settings = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev), sh_degree=3,
        campos=cam["campos"].to(dev), prefiltered=False, debug=False)
    params = {k: scene[k].to(dev).requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    means2D = torch.zeros(P, 3, device=dev, requires_grad=True)        # receives the densification proxy
    target = torch.rand(3, H, W, device=dev)

    for it in range(3):
        for t in list(params.values()) + [means2D]:
            t.grad = None
        out = SP.rasterize_tile_band(GaussianRasterizer, settings, rank, world, gather="fused" if world > 1 else "sync",
                                     means3D=params["means3D"], means2D=means2D, shs=params["shs"],
                                     opacities=params["opacities"], scales=params["scales"], rotations=params["rotations"])
        r0, r1 = out["band"]                                           # this rank's tile rows
        ys = slice(r0 * 16, min(H, r1 * 16))
        # a loss that only looks at this rank's rows; the other bands' rows of out["render"] are complete too
        # (the fused exchange has finished before rasterize_tile_band returns), e.g. for logging the full frame
        loss = (out["render"][:, ys] - target[:, ys]).abs().mean() + 0.05 * out["allmap"][6, ys].mean()
        loss.backward()                                                # gradients arrive summed over the ranks
        visible = out["radii"] > 0                                     # MAX over the ranks: visible in ANY band
        if rank == 0:
            print(f"iteration {it}: band loss {float(loss):.5f}, {int(visible.sum())} visible surfels, "
                  f"|dL/dmeans3D| = {float(params['means3D'].grad.norm()):.4e}")
    if world > 1:
        SP.release_symmetric_frames()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
