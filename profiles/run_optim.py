"""f3: the parameter update of one training iteration at 1 M surfels (59 floats per splat in 6 groups):
the reference's torch.optim.Adam(l, lr=0.0, eps=1e-15) (default foreach implementation), torch's own
fused=True variant for context, and diff_surfel_rasterization.optim.FusedAdam (one launch); plus the
densification statistics, eager (reference lines) vs fused.  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2d-gaussian-splatting_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

from diff_surfel_rasterization.optim import FusedAdam, densification_stats
from test_optim_gpu import _model

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda")


def time_ms(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {"workload": f"{P} surfels, 6 parameter groups, 59 floats per splat"}
for tag, make in (("torch_adam_foreach", lambda g: torch.optim.Adam(g, lr=0.0, eps=1e-15)),
                  ("torch_adam_fused", lambda g: torch.optim.Adam(g, lr=0.0, eps=1e-15, fused=True)),
                  ("surfel_fused_adam", lambda g: FusedAdam(g, lr=0.0, eps=1e-15))):
    params, groups = _model(P, dev, 3)
    for p in params.values():
        p.grad = torch.randn_like(p) * 1e-3
    opt = make(groups)
    out[tag + "_ms"] = time_ms(opt.step)
    del opt, params, groups
    torch.cuda.empty_cache()
nbytes = 59 * P * 4 * 7
out["surfel_fused_adam_GBps"] = nbytes / (out["surfel_fused_adam_ms"] * 1e-3) / 1e9
out["algorithmic_bytes"] = nbytes

radii = torch.randint(0, 40, (P,), device=dev, dtype=torch.int32) * (torch.rand(P, device=dev) > 0.3)
radii = radii.to(torch.int32)
grad = torch.randn(P, 3, device=dev)
accum, denom, maxr = torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, device=dev)


def eager():
    vis = radii > 0
    maxr[vis] = torch.max(maxr[vis], radii[vis])
    accum[vis] += torch.norm(grad[vis], dim=-1, keepdim=True)
    denom[vis] += 1


out["densify_stats_eager_ms"] = time_ms(eager)
out["densify_stats_fused_ms"] = time_ms(lambda: densification_stats(accum, denom, maxr, grad, radii))
print(json.dumps(out))
