"""What the host link gives on this box: pinned H2D alone, D2H alone and both at once (the e2e leg of
bench.py is bounded by these: 315 MB up + 331 MB down per step on the headline workload).

  python profiles/pcie_probe.py            # one JSON line
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def main():
    try:
        import bench
        bench.bind_to_gpu_numa_node(0)
    except Exception:
        pass
    dev = torch.device("cuda", 0)
    n = 256 << 20
    h_up, h_dn = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_up, d_dn = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed(up, dn, reps=8):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s_up.wait_event(a); s_dn.wait_event(a)
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s_up):
                    d_up.copy_(h_up, non_blocking=True)
            if dn:
                with torch.cuda.stream(s_dn):
                    h_dn.copy_(d_dn, non_blocking=True)
        e1, e2 = torch.cuda.Event(), torch.cuda.Event()
        e1.record(s_up); e2.record(s_dn)
        torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
        b.record()
        torch.cuda.synchronize()
        return reps * n / (a.elapsed_time(b) * 1e-3) / 1e9

    for _ in range(2):
        timed(True, True, 2)
    out = {"h2d_GBps": timed(True, False), "d2h_GBps": timed(False, True),
           "both_each_direction_GBps": timed(True, True), "buffer_MB": n >> 20}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
