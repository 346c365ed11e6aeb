"""Drop-in evidence that needs the reference checkout (build container only; skipped where
/root/reference does not exist, e.g. on the GPU box): the reference's OWN GaussianModel code —
training_setup, update_learning_rate, prune_points, densification_postfix, reset_opacity, capture/restore
— is run with `diff_surfel_rasterization.optim.FusedAdam` substituted for torch.optim.Adam, and must leave
the optimizer in exactly the state it leaves a torch.optim.Adam in.  (FusedAdam.step itself needs a GPU and is
covered by tests/test_optim_gpu.py; here the per-parameter state is planted by hand, as Adam would create it.)
Runs in a subprocess because the reference hard-codes device="cuda" and torch has to be patched to CPU."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

SCRIPT = textwrap.dedent('''
    import sys, types
    import torch
    sys.path.insert(0, {root!r} + "/tests/golden"); sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/2d-gaussian-splatting_b200")
    from diff_surfel_rasterization.optim import FusedAdam           # the real package, before the stubs go in
    import make_golden as MG
    MG.cpu_patches(); MG.stub_modules({{}})                           # plyfile / simple_knn / ... stubs, torch -> CPU
    sys.path.insert(0, {ref!r})
    import scene.gaussian_model as GM

    def build(adam_cls):
        torch.optim.Adam, keep = adam_cls, torch.optim.Adam          # training_setup calls torch.optim.Adam(l, lr=0.0, eps=1e-15)
        try:
            g = torch.Generator("cpu").manual_seed(5)
            P = 12
            pc = GM.GaussianModel(3)
            pc._xyz = torch.nn.Parameter(torch.randn(P, 3, generator=g))
            pc._features_dc = torch.nn.Parameter(torch.randn(P, 1, 3, generator=g))
            pc._features_rest = torch.nn.Parameter(torch.randn(P, 15, 3, generator=g))
            pc._opacity = torch.nn.Parameter(torch.randn(P, 1, generator=g))
            pc._scaling = torch.nn.Parameter(torch.randn(P, 2, generator=g))
            pc._rotation = torch.nn.Parameter(torch.randn(P, 4, generator=g))
            pc.max_radii2D = torch.zeros(P)
            pc.spatial_lr_scale = 5.0
            args = types.SimpleNamespace(percent_dense=0.01, position_lr_init=0.00016, position_lr_final=0.0000016,
                                         position_lr_delay_mult=0.01, position_lr_max_steps=30000, feature_lr=0.0025,
                                         opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001)
            pc.training_setup(args)
        finally:
            torch.optim.Adam = keep
        opt = pc.optimizer
        for grp in opt.param_groups:                                   # state as Adam creates it on the first step
            p = grp["params"][0]
            opt.state[p] = {{"step": torch.tensor(3.0), "exp_avg": torch.randn(p.shape, generator=g),
                            "exp_avg_sq": torch.rand(p.shape, generator=g)}}
        lr = pc.update_learning_rate(1000)
        mask = torch.zeros(12, dtype=torch.bool); mask[[1, 4, 9]] = True
        pc.prune_points(mask)
        n = 4
        pc.densification_postfix(torch.ones(n, 3), torch.ones(n, 1, 3), torch.ones(n, 15, 3), torch.ones(n, 1),
                                 torch.ones(n, 2), torch.ones(n, 4))
        pc.reset_opacity()
        snap = pc.capture()                                            # checkpoint path: optimizer.state_dict()
        return pc, lr, snap

    a, lr_a, snap_a = build(torch.optim.Adam)
    b, lr_b, snap_b = build(FusedAdam)
    assert isinstance(b.optimizer, FusedAdam) and isinstance(b.optimizer, torch.optim.Adam)
    assert lr_a == lr_b
    for ga, gb in zip(a.optimizer.param_groups, b.optimizer.param_groups):
        assert ga["name"] == gb["name"] and ga["lr"] == gb["lr"] and ga["eps"] == gb["eps"] == 1e-15 and ga["betas"] == gb["betas"]
        pa, pb = ga["params"][0], gb["params"][0]
        assert pa.shape == pb.shape and pa.shape[0] == 12 - 3 + 4 and torch.equal(pa, pb)
        sa, sb = a.optimizer.state[pa], b.optimizer.state[pb]
        assert set(sa) == set(sb) == {{"step", "exp_avg", "exp_avg_sq"}}
        for k in sa:
            assert torch.equal(sa[k], sb[k]), (ga["name"], k)
    sd_a, sd_b = snap_a[-2], snap_b[-2]                                # optimizer.state_dict() inside capture()
    assert sd_a["param_groups"][0].keys() == sd_b["param_groups"][0].keys() or set(sd_a["param_groups"][0]) <= set(sd_b["param_groups"][0])
    b.optimizer.load_state_dict(sd_a)                                  # a checkpoint written with torch's Adam loads
    print("INTEROP_OK", len(a.optimizer.param_groups))
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (build container only)")
def test_reference_densification_code_runs_unchanged_on_fused_adam():
    r = subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT, ref=REF)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "INTEROP_OK 6" in r.stdout
