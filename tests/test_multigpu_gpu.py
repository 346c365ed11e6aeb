"""2-GPU NCCL test of the tile-band partition (run with `gpurun --gpus 2`; skipped on 1-GPU boxes)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "2d-gaussian-splatting_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import surfel_parallel as SP
    import surfel_scenes as S
    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        W, H, P = 640, 360, 20000
        cam = S.make_camera(W, H)
        scene = S.make_scene(P, W, H, 31)
        rs = GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=torch.zeros(3, device=dev),
            scale_modifier=1.0, viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev), sh_degree=3,
            campos=cam["campos"].to(dev), prefiltered=False, debug=False)
        leaf = {k: v.to(dev).requires_grad_(True) for k, v in scene.items()}
        m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
        res = SP.rasterize_tile_band(GaussianRasterizer, rs, rank, world, means3D=leaf["means3D"], means2D=m2d,
                                     shs=leaf["shs"], opacities=leaf["opacities"], scales=leaf["scales"], rotations=leaf["rotations"])
        gc, go = S.make_cotangents(W, H, 31)
        ((res["render"] * gc.to(dev)).sum() + (res["allmap"] * go.to(dev)).sum()).backward()
        grads = [leaf[k].grad for k in ("means3D", "scales", "rotations", "opacities", "shs")]   # already summed over the ranks
        # the same frame with the exchange done by the render kernel's own stores (symmetric memory: NVSwitch
        # multicast if the group has it, and peer stores), and with asynchronous gathers
        variants = {}
        for mode in ("fused", "fused_multicast", "async"):
            with torch.no_grad():
                rv = SP.rasterize_tile_band(GaussianRasterizer, rs, rank, world, gather=mode, means3D=leaf["means3D"], means2D=m2d,
                                            shs=leaf["shs"], opacities=leaf["opacities"], scales=leaf["scales"],
                                            rotations=leaf["rotations"])
            rv["wait"]()
            variants[mode] = (rv["render"].clone(), rv["allmap"].clone(), SP._last.get("fused_via"))
        if rank == 0:
            # single-GPU reference on the same device
            leaf2 = {k: v.to(dev).requires_grad_(True) for k, v in scene.items()}
            m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
            color, radii, allmap = GaussianRasterizer(rs)(means3D=leaf2["means3D"], means2D=m2, shs=leaf2["shs"],
                                                          opacities=leaf2["opacities"], scales=leaf2["scales"], rotations=leaf2["rotations"])
            ((color * gc.to(dev)).sum() + (allmap * go.to(dev)).sum()).backward()
            ok_img = torch.equal(res["render"], color) and torch.equal(res["allmap"], allmap) and torch.equal(res["radii"], radii)
            errs = {"means2D": float((m2d.grad - m2.grad).abs().max() / (m2.grad.abs().max() + 1e-30))}
            for k, g in zip(("means3D", "scales", "rotations", "opacities", "shs"), grads):
                ref = leaf2[k].grad
                errs[k] = float((g - ref).abs().max() / (ref.abs().max() + 1e-30))
            ok_var = {m: bool(torch.equal(v[0], color) and torch.equal(v[1], allmap)) for m, v in variants.items()}
            torch.save({"ok_img": ok_img, "errs": errs, "ok_var": ok_var, "via": {m: v[2] for m, v in variants.items()}}, out)
    finally:
        dist.destroy_process_group()


def test_tile_band_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["ok_img"], "the frame completed in place (or the MAX-reduced radii) differs from the single-GPU result"
    assert max(r["errs"].values()) < 1e-3, r["errs"]
    print("exchange variants:", r["ok_var"], r["via"])
    assert all(r["ok_var"].values()), f"a frame completed by a fused / asynchronous exchange differs from the single-GPU result: {r['ok_var']}"
