"""Host-side anatomy of one tile-band step (torchrun, N ranks): where the host spends its time between the
marks of the autograd node and of surfel_parallel._BandFrame, and how many cudaMalloc / cudaFree calls the
caching allocator makes per step.  Written to find why the band forward took 14-30 ms of device-idle time at
N=2 when its kernels sum to 2 ms.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 profiles/band_probe.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-gaussian-splatting_b200")]


def main():
    import torch
    import torch.distributed as dist
    import diff_surfel_rasterization as dsr
    import surfel_parallel as SP
    import surfel_scenes as S
    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    workload = os.environ.get("PROBE_WORKLOAD", "config5")
    P, W, H = S.CONFIGS[workload]
    scene, cam = S.named(workload)
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev), sh_degree=3,
        campos=cam["campos"].to(dev), prefiltered=False, debug=False)
    names = ("means3D", "scales", "rotations", "opacities", "shs")
    leaf = {k: scene[k].to(dev).requires_grad_(True) for k in names}
    m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
    gc, go = S.make_cotangents(W, H, 5)
    gc, go = gc.to(dev), go.to(dev)

    def step():
        for t in list(leaf.values()) + [m2d]:
            t.grad = None
        res = SP.rasterize_tile_band(GaussianRasterizer, rs, rank, world, means3D=leaf["means3D"], means2D=m2d,
                                     shs=leaf["shs"], opacities=leaf["opacities"], scales=leaf["scales"],
                                     rotations=leaf["rotations"])
        torch.autograd.backward([res["render"], res["allmap"]], [gc, go])
        return res

    for _ in range(3):
        step(); torch.cuda.synchronize()
    dist.barrier(); torch.cuda.synchronize()
    rows = []
    for it in range(5):
        st0 = torch.cuda.memory_stats(dev)
        dsr.trace_host(True)
        t0 = time.perf_counter_ns()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        t1 = time.perf_counter_ns()
        torch.cuda.synchronize()
        t2 = time.perf_counter_ns()
        marks = dsr.trace_host(False)
        st1 = torch.cuda.memory_stats(dev)
        seq, prev = [], t0
        for tag, t, _ in marks:
            seq.append((tag, round((t - prev) / 1e6, 3))); prev = t
        seq.append(("step_returned", round((t1 - prev) / 1e6, 3)))
        rows.append({"iter": it, "host_ms_to_return": round((t1 - t0) / 1e6, 3), "host_ms_to_drain": round((t2 - t0) / 1e6, 3),
                     "device_ms": round(e0.elapsed_time(e1), 3),
                     "cudaMalloc": st1["num_device_alloc"] - st0["num_device_alloc"],
                     "cudaFree": st1["num_device_free"] - st0["num_device_free"],
                     "reserved_GB": round(st1["reserved_bytes.all.current"] / 1e9, 2),
                     "ms_since_previous_mark": seq})
    if rank == 0:
        print(json.dumps({"world": world, "workload": workload,
                          "env": {k: os.environ.get(k) for k in ("TORCH_NCCL_AVOID_RECORD_STREAMS", "PYTORCH_CUDA_ALLOC_CONF")},
                          "steps": rows}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
