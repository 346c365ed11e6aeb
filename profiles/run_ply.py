"""f4: a 1 M-splat model file -> rasterizer inputs.  (a) the reference's way: numpy column copies on the
CPU + six host->device tensor constructions (load_ply restated in tests/test_ply_gpu.py) + the getters'
activations; (b) surfel_ply.load_ply: raw rows -> pinned -> device -> ONE unpack kernel; and the unpack /
pack kernels alone with the rows already resident.  Prints one JSON line."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2d-gaussian-splatting_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

import surfel_ply as PLY
from test_ply_cpu import random_model, reference_file_bytes
from test_ply_gpu import reference_load

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda")
blob = reference_file_bytes(*random_model(P, 1))
path = os.path.join(tempfile.mkdtemp(), "point_cloud.ply")
open(path, "wb").write(blob)
out = {"workload": f"{P} splats, {len(blob) / 1e6:.0f} MB PLY"}

torch.cuda.synchronize()
t0 = time.perf_counter()
ref = reference_load(open(path, "rb").read())
ref = {k: v.to(dev) for k, v in ref.items()}
act = (torch.sigmoid(ref["opacity"]), torch.exp(ref["scaling"]), torch.nn.functional.normalize(ref["rotation"]),
       torch.cat((ref["features_dc"], ref["features_rest"]), dim=1))
torch.cuda.synchronize()
out["reference_way_ms"] = (time.perf_counter() - t0) * 1e3

PLY.load_ply(path)                                           # warm-up (page cache, allocator)
torch.cuda.synchronize()
t0 = time.perf_counter()
m = PLY.load_ply(path)
torch.cuda.synchronize()
out["load_ply_ms"] = (time.perf_counter() - t0) * 1e3

count, names, offset = PLY.parse_header(blob[:1 << 16])
rows = torch.frombuffer(bytearray(blob[offset:]), dtype=torch.float32).reshape(count, len(names)).to(dev)


def time_ms(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out["unpack_kernel_ms"] = time_ms(lambda: PLY.unpack_rows(rows, names, True))
out["unpack_GBps"] = P * (61 + 58) * 4 / (out["unpack_kernel_ms"] * 1e-3) / 1e9
raw = PLY.unpack_rows(rows, names, False)
args = (raw["means3D"], raw["shs"][:, :1].contiguous(), raw["shs"][:, 1:].contiguous(), raw["opacities"], raw["scales"], raw["rotations"])
out["pack_kernel_ms"] = time_ms(lambda: PLY.pack_rows(*args))
out["same_result"] = bool(torch.equal(m["shs"], act[3]) and torch.allclose(m["opacities"], act[0], rtol=2e-6, atol=1e-7))
print(json.dumps(out))
