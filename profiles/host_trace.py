"""Where does the host spend a step?  (diagnostic for the multi-process step-time gap)

Runs the headline fwd+bwd loop with the autograd node's host trace switched on
(`diff_surfel_rasterization.trace_host`) and prints, per step and as medians, the two critical sections:

  window A  "R is known" -> "backward is launched"        must fit in the queued forward work (~0.57 ms)
  window B  "backward is launched" -> "next preprocess is launched"   must fit in the backward (~1.05 ms)

plus which thread ran forward / backward (autograd hands the backward to a device thread unless
torch.autograd.set_multithreading_enabled(False)), the wall time per step and the CUDA-event time per step.
Launch it like bench.py: plain `python profiles/host_trace.py`, or under torchrun for N ranks (every rank
prints its own JSON line to stderr, rank 0 to stdout).

  python profiles/host_trace.py [--steps 50] [--inline-backward] [--init-nccl]
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "2d-gaussian-splatting_b200"))
import torch

import bench
import diff_surfel_rasterization as dsr
import surfel_scenes as S
from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--inline-backward", action="store_true", help="torch.autograd.set_multithreading_enabled(False)")
    ap.add_argument("--init-nccl", action="store_true", help="create an NCCL process group even for one rank")
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    bench.bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 or args.init_nccl:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29555")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        dist.barrier()
    if args.inline_backward:
        torch.autograd.set_multithreading_enabled(False)
    P, W, H = S.CONFIGS["headline"]
    scene, cam = S.named("headline")
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev), sh_degree=3,
        campos=cam["campos"].to(dev), prefiltered=False, debug=False)
    rast = GaussianRasterizer(rs)
    leaf = {k: v.to(dev).requires_grad_(True) for k, v in scene.items()}
    m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
    gc, go = (t.to(dev) for t in S.make_cotangents(W, H, 5))

    def step():
        for t in list(leaf.values()) + [m2d]:
            t.grad = None
        color, radii, allmap = rast(means3D=leaf["means3D"], means2D=m2d, shs=leaf["shs"], opacities=leaf["opacities"],
                                    scales=leaf["scales"], rotations=leaf["rotations"])
        torch.autograd.backward([color, allmap], [gc, go])

    for _ in range(10):
        step()
    torch.cuda.synchronize()
    dsr.trace_host(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    marks = dsr.trace_host(False)
    # split into steps at every fwd_enter
    steps, cur = [], None
    for tag, t, tid in marks:
        if tag == "fwd_enter":
            cur = {}
            steps.append(cur)
        cur[tag] = (t, tid)
    us = lambda a, b: (b[0] - a[0]) / 1e3
    A = [us(s["R_known"], s["bwd_launched"]) for s in steps if "bwd_launched" in s]
    B = [us(steps[i]["bwd_launched"], steps[i + 1]["preprocess_launched"]) for i in range(len(steps) - 1)]
    wait = [us(s["speculative_work_launched"], s["R_known"]) for s in steps]
    hop = [us(s["fwd_exit"], s["bwd_enter"]) for s in steps if "bwd_enter" in s]
    wall = [us(steps[i]["fwd_enter"], steps[i + 1]["fwd_enter"]) for i in range(len(steps) - 1)]
    med = lambda v: round(statistics.median(v), 1)
    p95 = lambda v: round(sorted(v)[int(0.95 * (len(v) - 1))], 1)
    out = {"rank": rank, "world": world, "inline_backward": args.inline_backward, "steps": len(steps),
           "gpu_ms_per_step": e0.elapsed_time(e1) / args.steps,
           "wall_us_per_step": {"median": med(wall), "p95": p95(wall)},
           "window_A_R_known_to_bwd_launched_us": {"median": med(A), "p95": p95(A), "budget": "~570 (queued forward work)"},
           "window_B_bwd_launched_to_next_preprocess_us": {"median": med(B), "p95": p95(B), "budget": "~1050 (backward)"},
           "host_blocked_waiting_for_R_us": {"median": med(wait), "p95": p95(wait)},
           "fwd_exit_to_bwd_enter_us": {"median": med(hop), "p95": p95(hop)},
           "backward_on_other_thread": steps[0]["fwd_enter"][1] != steps[0]["bwd_enter"][1],
           "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")}
    print(json.dumps(out), file=sys.stdout if rank == 0 else sys.stderr)
    if world > 1 or args.init_nccl:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
