"""Why the host-buffer (e2e) leg does not scale 1:1 with the number of ranks (efficiency 0.87 / 0.46
/ 0.51 at N = 2 / 4 / 8): every rank moves ~315 MB up and ~331 MB down per step over its own PCIe link, but
the links of the GPUs of one socket end in the same memory controllers.  This probe runs N ranks that copy
pinned buffers in both directions AT THE SAME TIME and reports per-rank and aggregate GB/s for three
placements of the pinned buffers:
    local       pages on the NUMA node the rank's GPU hangs off (what bench.py does)
    interleave  pages interleaved over all NUMA nodes (set_mempolicy(MPOL_INTERLEAVE) before the allocation)
    remote      pages on another node (every transfer crosses the socket interconnect)
and the same for one rank alone.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
        profiles/pcie_concurrent.py
Rank 0 prints one JSON line."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

MPOL_DEFAULT, MPOL_BIND, MPOL_INTERLEAVE = 0, 2, 3


def numa_nodes():
    try:
        return sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except OSError:
        return [0]


def set_mempolicy(mode, nodes):
    mask = ctypes.c_ulong(sum(1 << n for n in nodes))
    libc = ctypes.CDLL(None, use_errno=True)
    r = libc.syscall(238, ctypes.c_int(mode), ctypes.byref(mask), ctypes.c_ulong(64))   # __NR_set_mempolicy (x86_64)
    return r == 0


def gpu_node(index):
    import subprocess
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        return int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
    except Exception:
        return -1


def measure(dev, n, seconds=0.6):
    h_up, h_dn = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    h_up.fill_(1); h_dn.fill_(2)                                       # touch every page under the policy in force
    d_up, d_dn = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        s_up.wait_event(a); s_dn.wait_event(a)
        for _ in range(reps):
            with torch.cuda.stream(s_up):
                d_up.copy_(h_up, non_blocking=True)
            with torch.cuda.stream(s_dn):
                h_dn.copy_(d_dn, non_blocking=True)
        e1, e2 = torch.cuda.Event(), torch.cuda.Event()
        e1.record(s_up); e2.record(s_dn)
        torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3
    run(2)
    reps = max(4, int(seconds / max(run(2) / 2, 1e-4)))
    if dist.is_initialized():
        t = torch.tensor([reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN); reps = int(t.item())
        dist.barrier()
    dt = run(reps)
    return reps * n / dt / 1e9            # GB/s per direction (both directions run concurrently)


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nodes = numa_nodes()
    mine = gpu_node(local)
    n = 256 << 20
    res = {}
    for name in ("local", "interleave", "remote"):
        if name == "local":
            ok = set_mempolicy(MPOL_BIND, [mine]) if mine >= 0 else False
        elif name == "interleave":
            ok = set_mempolicy(MPOL_INTERLEAVE, nodes)
        else:
            others = [x for x in nodes if x != mine] or nodes
            ok = set_mempolicy(MPOL_BIND, [others[0]])
        v = measure(dev, n)
        set_mempolicy(MPOL_DEFAULT, [])
        t = torch.tensor([v], device=dev)
        if world > 1:
            g = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(g, t)
            vals = [float(x) for x in g]
        else:
            vals = [v]
        res[name] = {"policy_applied": bool(ok), "per_rank_GBps_each_direction": [round(x, 1) for x in vals],
                     "aggregate_GBps_each_direction": round(sum(vals), 1)}
    gn = torch.tensor([mine], device=dev)
    if world > 1:
        g = [torch.zeros_like(gn) for _ in range(world)]
        dist.all_gather(g, gn)
        gnodes = [int(x) for x in g]
    else:
        gnodes = [mine]
    if rank == 0:
        print(json.dumps({"ranks": world, "numa_nodes": nodes, "gpu_numa_node_per_rank": gnodes, "buffer_MB": n >> 20,
                          "what": "both directions at once, every rank at once; GB/s per direction", "placements": res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
