"""N>1 host-side logic on CPU: world_size-2 gloo processes (no GPU).

The per-rank renderer here is the CPU oracle (tests may use it); what is under test is the sharding
code in surfel_parallel.py — view assignment, tile-row bands, the band all-gather, cotangent slicing
and the gradient all-reduce — i.e. everything the N-GPU path adds around the CUDA op.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "2d-gaussian-splatting_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import surfel_parallel as SP
    import surfel_scenes as S
    from oracle import surfel_oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, H, P = 112, 100, 500          # 7 tile rows: uneven bands (3 + 4)
        cam = S.make_camera(W, H)
        scene = S.make_scene(P, W, H, 21, depth_complexity=25)
        sn, cn = S.to_numpy(scene), S.to_numpy(cam)
        bg = np.array([0.1, 0.3, 0.2], np.float32)
        band = SP.tile_row_band(H, rank, world)
        pre, binned, img = O.forward(sn, cn, bg, row0=band[0], row1=band[1])
        planes = torch.from_numpy(np.concatenate([img["color"], img["others"]], 0))
        s, e = SP.band_pixel_rows(H, band)
        planes[:, :s] = 0; planes[:, e:] = 0      # what the CUDA op leaves outside its band
        full = SP.gather_band_outputs(planes, H, rank, world)
        # backward: slice the cotangent, run the band backward, all-reduce per-splat gradients
        gc, go = S.make_cotangents(W, H, 21)
        g = SP.slice_band_cotangent(torch.cat([gc, go], 0), H, rank, world)
        grads = O.backward(sn, cn, bg, pre, binned, img, g[:3].numpy(), g[3:].numpy())
        keys = ["dL_dmeans3D", "dL_dscales", "dL_drotations", "dL_dopacity", "dL_dshs"]
        # the densification proxy is not additive over bands by construction? it is linear in dL_dT: include it
        tens = [torch.from_numpy(grads[k].astype(np.float32)) for k in keys + ["dL_dmeans2D"]]
        SP.allreduce_gradients(tens)
        # copy-free exchange: padded frame, equal bands, in-place all-gather per plane
        eb = SP.equal_band(H, rank, world)
        pre_e, bin_e, img_e = O.forward(sn, cn, bg, row0=eb[0], row1=eb[1])
        buf = SP.padded_frame(10, H, W, world, "cpu")
        buf.fill_(float("nan"))
        es, ee = SP.band_pixel_rows(H, eb)
        buf[:3, es:ee] = torch.from_numpy(img_e["color"][:, es:ee]); buf[3:, es:ee] = torch.from_numpy(img_e["others"][:, es:ee])
        SP.allgather_frame_inplace(buf, H, rank, world)
        buf2 = SP.padded_frame(10, H, W, world, "cpu")
        buf2.fill_(float("nan"))
        buf2[:, es:ee] = buf[:, es:ee]
        works = SP.allgather_frame_inplace(buf2, H, rank, world, async_op=True)     # enqueue only ...
        assert len(works) == 10
        for w in works:                                                              # ... complete on wait()
            w.wait()
        assert torch.equal(buf2[:, :H], buf[:, :H])
        views = SP.shard_views(5, rank, world)
        if rank == 0:
            torch.save({"full": full, "grads": dict(zip(keys + ["dL_dmeans2D"], tens)), "views": views, "band": band,
                        "frame_inplace": buf[:, :H].clone(), "equal_band": eb}, out)
        else:
            torch.save({"views": views, "band": band}, out + ".r1")
    finally:
        dist.destroy_process_group()


def test_tile_band_and_view_sharding_world2(oracle, tmp_path):
    import surfel_parallel as SP
    import surfel_scenes as S
    out = str(tmp_path / "r0.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = torch.load(out), torch.load(out + ".r1")
    assert r0["views"] == [0, 2, 4] and r1["views"] == [1, 3]
    assert r0["band"] == (0, 3) and r1["band"] == (3, 7)
    W, H, P = 112, 100, 500
    cam = S.make_camera(W, H)
    scene = S.make_scene(P, W, H, 21, depth_complexity=25)
    sn, cn = S.to_numpy(scene), S.to_numpy(cam)
    bg = np.array([0.1, 0.3, 0.2], np.float32)
    pre, binned, img = oracle.forward(sn, cn, bg)
    ref = np.concatenate([img["color"], img["others"]], 0)
    np.testing.assert_array_equal(r0["full"].numpy(), ref)          # stitched frame is bit-identical
    np.testing.assert_array_equal(r0["frame_inplace"].numpy(), ref)  # and so is the frame completed in place
    assert r0["equal_band"] == (0, 4)
    gc, go = S.make_cotangents(W, H, 21)
    full = oracle.backward(sn, cn, bg, pre, binned, img, gc.numpy(), go.numpy())
    for k, v in r0["grads"].items():
        np.testing.assert_allclose(v.numpy().reshape(full[k].shape), full[k], rtol=2e-4,
                                   atol=2e-4 * np.abs(full[k]).max(), err_msg=k)


def test_band_helpers():
    import surfel_parallel as SP
    for H in (16, 100, 1080, 4320):
        gy = SP.tile_rows(H)
        for world in (1, 2, 3, 8):
            if world > gy:
                continue
            bands = [SP.tile_row_band(H, r, world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == gy
            assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
            assert max(b[1] - b[0] for b in bands) - min(b[1] - b[0] for b in bands) <= 1
    assert SP.shard_views(8, 3, 8) == [3] and SP.shard_views(3, 5, 8) == []
    for H, world in ((4320, 8), (100, 2), (100, 8), (16, 3)):
        gy, rp = SP.tile_rows(H), SP.equal_band_rows(H, world)
        eb = [SP.equal_band(H, r, world) for r in range(world)]
        assert eb[0][0] == 0 and eb[-1][1] == gy and all(eb[i][1] == eb[i + 1][0] for i in range(world - 1))
        assert all(b[1] - b[0] <= rp for b in eb) and SP.padded_frame(1, H, 8, world, "cpu").shape[1] == rp * world * 16
    x = torch.arange(2 * 40 * 8, dtype=torch.float32).reshape(2, 40, 8)
    assert torch.equal(SP.gather_band_outputs(x, 40, 0, 1), x)


def test_rasterize_tile_band_rejects_unknown_modes():
    """Argument validation happens before anything touches a device."""
    import surfel_parallel as SP
    with pytest.raises(ValueError):
        SP.rasterize_tile_band(None, None, 0, 1, grad_reduce="sum")
    with pytest.raises(ValueError):
        SP.rasterize_tile_band(None, None, 0, 1, gather="nccl")
    with pytest.raises(Exception, match="excatly one"):
        SP.rasterize_tile_band(None, None, 0, 1, means3D=torch.zeros(1, 3), means2D=torch.zeros(1, 3), opacities=torch.zeros(1, 1))


def test_tile_band_example_compiles():
    import py_compile
    py_compile.compile(os.path.join(ROOT, "examples", "tile_band_step.py"), doraise=True)

