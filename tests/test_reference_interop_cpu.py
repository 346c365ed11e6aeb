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


RENDER_SCRIPT = textwrap.dedent('''
    import sys, types
    import numpy as np
    import torch
    sys.path.insert(0, {root!r} + "/tests/golden"); sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/2d-gaussian-splatting_b200")
    import diff_surfel_rasterization as real                       # OUR package: same import name as upstream's
    import surfel_scenes as S
    import make_golden as MG
    MG.cpu_patches(); MG.stub_modules({{}})
    sys.modules["diff_surfel_rasterization"] = real                # undo the rasterizer stub: the reference must import ours
    calls = []

    def fake_native(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
        calls.append(dict(means3D=means3D, means2D=means2D, sh=sh, colors_precomp=colors_precomp, opacities=opacities,
                          scales=scales, rotations=rotations, cov3Ds_precomp=cov3Ds_precomp, rs=rs))
        H, W = rs.image_height, rs.image_width
        return torch.zeros(3, H, W), torch.ones(means3D.shape[0], dtype=torch.int32), torch.ones(7, H, W)
    real.rasterize_gaussians = fake_native                         # everything ABOVE the native call is the real code
    sys.path.insert(0, {ref!r})
    from gaussian_renderer import render                           # /root/reference/gaussian_renderer/__init__.py:14 imports ours
    from scene.cameras import Camera
    from scene.gaussian_model import GaussianModel

    W, H, P = 64, 48, 10
    Rm, tv = S.look_at_rotation(10, 5), np.array([0.1, 0.0, 0.3])
    mycam = S.make_camera(W, H, R=Rm, t=tv)
    cam = Camera(colmap_id=0, R=Rm, T=tv, FoVx=mycam["FoVx"], FoVy=mycam["FoVy"], image=torch.zeros(3, H, W),
                 gt_alpha_mask=None, image_name="g", uid=0, data_device="cpu")
    scene = S.make_scene(P, W, H, 4, depth_complexity=2)
    pc = GaussianModel(3); pc.active_sh_degree = 3
    pc._xyz, pc._scaling, pc._rotation = scene["means3D"], torch.log(scene["scales"]), scene["rotations"]
    pc._opacity = torch.log(scene["opacities"] / (1 - scene["opacities"]))
    pc._features_dc, pc._features_rest = scene["shs"][:, :1].contiguous(), scene["shs"][:, 1:].contiguous()
    for cov_py in (False, True):
        for sh_py in (False, True):
            # render() forces pipe.convert_SHs_python = False (gaussian_renderer/__init__.py:82): its colors_precomp
            # path is reached through override_color
            pipe = types.SimpleNamespace(compute_cov3D_python=cov_py, convert_SHs_python=False, depth_ratio=0.0, debug=False)
            rets = render(cam, pc, pipe, torch.zeros(3), override_color=torch.rand(P, 3) if sh_py else None)
            c = calls[-1]
            rs = c["rs"]
            assert isinstance(rs, real.GaussianRasterizationSettings)
            assert (rs.image_height, rs.image_width, rs.sh_degree, rs.prefiltered, rs.scale_modifier) == (H, W, 3, False, 1.0)
            assert isinstance(rs.tanfovx, float) and rs.viewmatrix.shape == (4, 4) and rs.projmatrix.shape == (4, 4) and rs.campos.shape == (3,)
            assert c["means3D"].shape == (P, 3) and c["means2D"].shape == (P, 3) and c["opacities"].shape == (P, 1)
            assert (c["cov3Ds_precomp"].numel() > 0) == cov_py and (c["scales"].numel() > 0) == (not cov_py) and (c["rotations"].numel() > 0) == (not cov_py)
            assert (c["colors_precomp"].numel() > 0) == sh_py and (c["sh"].numel() > 0) == (not sh_py)
            if cov_py:
                assert c["cov3Ds_precomp"].shape == (P, 9)
            if not sh_py:
                assert c["sh"].shape == (P, 16, 3)
            assert set(rets) >= {{"render", "viewspace_points", "visibility_filter", "radii", "rend_alpha", "rend_normal",
                                 "rend_dist", "surf_depth", "surf_normal"}}
            assert rets["visibility_filter"].dtype == torch.bool and rets["visibility_filter"].all()
    print("RENDER_INTEROP_OK", len(calls))
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (build container only)")
def test_reference_render_drives_our_python_surface():
    """The reference's unmodified gaussian_renderer.render() imports OUR diff_surfel_rasterization, builds
    GaussianRasterizationSettings with its own keywords and calls GaussianRasterizer with its own argument patterns
    (compute_cov3D_python on/off x SH / override_color inputs); only the native call below the Python surface
    is replaced (no GPU here)."""
    r = subprocess.run([sys.executable, "-c", RENDER_SCRIPT.format(root=ROOT, ref=REF)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "RENDER_INTEROP_OK 4" in r.stdout
