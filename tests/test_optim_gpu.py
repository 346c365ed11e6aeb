"""fused Adam step + densification statistics against the reference's own calls
(torch.optim.Adam(l, lr=0.0, eps=1e-15) as built at /root/reference/scene/gaussian_model.py:148-166;
train.py:125-128 + gaussian_model.py:405-407 for the statistics)."""
import pytest
import torch

gpu = pytest.mark.gpu


def _model(P, dev, seed):
    g = torch.Generator("cpu").manual_seed(seed)
    shapes = {"xyz": (P, 3), "f_dc": (P, 1, 3), "f_rest": (P, 15, 3), "opacity": (P, 1), "scaling": (P, 2), "rotation": (P, 4)}
    lrs = {"xyz": 1.6e-4 * 5.0, "f_dc": 2.5e-3, "f_rest": 2.5e-3 / 20.0, "opacity": 0.05, "scaling": 0.005, "rotation": 0.001}
    params = {k: torch.nn.Parameter(torch.randn(s, generator=g).to(dev)) for k, s in shapes.items()}
    groups = [{"params": [params[k]], "lr": lrs[k], "name": k} for k in shapes]
    return params, groups


@gpu
@pytest.mark.parametrize("P", [1, 1003, 50000])
def test_fused_adam_matches_torch_adam(P):
    from diff_surfel_rasterization.optim import FusedAdam
    dev = torch.device("cuda")
    pa, ga = _model(P, dev, 3)
    pb, gb = _model(P, dev, 3)
    ref = torch.optim.Adam(ga, lr=0.0, eps=1e-15)
    fused = FusedAdam(gb, lr=0.0, eps=1e-15)
    g = torch.Generator("cpu").manual_seed(11)
    for it in range(6):
        for k in pa:
            grad = torch.randn(pa[k].shape, generator=g) * (10.0 ** torch.randint(-6, 1, (1,), generator=g).item())
            if it == 2 and k == "opacity":
                grad.zero_()                                   # exact zeros: m, v stay finite with eps = 1e-15
            pa[k].grad = grad.to(dev)
            pb[k].grad = grad.to(dev).clone()
        if it == 4:                                            # the xyz learning-rate schedule rewrites group["lr"]
            for opt in (ref, fused):
                for grp in opt.param_groups:
                    if grp["name"] == "xyz":
                        grp["lr"] = 3.1e-5
        ref.step(); fused.step()
        ref.zero_grad(set_to_none=True); fused.zero_grad(set_to_none=True)
    for k in pa:
        torch.testing.assert_close(pb[k].data, pa[k].data, rtol=2e-6, atol=2e-7)
        sa, sb = ref.state[pa[k]], fused.state[pb[k]]
        assert float(sa["step"]) == float(sb["step"]) == 6.0
        torch.testing.assert_close(sb["exp_avg"], sa["exp_avg"], rtol=5e-6, atol=1e-12)
        torch.testing.assert_close(sb["exp_avg_sq"], sa["exp_avg_sq"], rtol=5e-6, atol=1e-18)
    # same state_dict layout: the reference's densification code edits it in place
    assert set(fused.state_dict()["state"][0].keys()) == set(ref.state_dict()["state"][0].keys())


@gpu
def test_fused_adam_survives_state_surgery():
    """cat_tensors_to_optimizer of the reference (gaussian_model.py:303-324) replaces a parameter and
    extends its state; the fused step must pick the new tensors up."""
    from diff_surfel_rasterization.optim import FusedAdam
    dev = torch.device("cuda")
    res = []
    for cls in (torch.optim.Adam, FusedAdam):
        p = torch.nn.Parameter(torch.linspace(-1, 1, 30, device=dev).reshape(10, 3).clone())
        opt = cls([{"params": [p], "lr": 0.01, "name": "xyz"}], lr=0.0, eps=1e-15)
        p.grad = torch.full_like(p, 0.5); opt.step()
        group = opt.param_groups[0]
        stored = opt.state.get(group["params"][0])
        ext = torch.ones(4, 3, device=dev)
        stored["exp_avg"] = torch.cat((stored["exp_avg"], torch.zeros_like(ext)), dim=0)
        stored["exp_avg_sq"] = torch.cat((stored["exp_avg_sq"], torch.zeros_like(ext)), dim=0)
        del opt.state[group["params"][0]]
        group["params"][0] = torch.nn.Parameter(torch.cat((group["params"][0], ext), dim=0).requires_grad_(True))
        opt.state[group["params"][0]] = stored
        q = group["params"][0]
        q.grad = torch.full_like(q, -0.25); opt.step()
        res.append(q.detach().clone())
    torch.testing.assert_close(res[1], res[0], rtol=2e-6, atol=1e-7)


@gpu
@pytest.mark.parametrize("P", [1, 777, 100000])
def test_densification_stats(P):
    from diff_surfel_rasterization.optim import densification_stats
    dev = torch.device("cuda")
    g = torch.Generator("cpu").manual_seed(5)
    radii = (torch.randint(-2, 40, (P,), generator=g).clamp_min(0)).to(torch.int32).to(dev)
    grad = torch.randn(P, 3, generator=g).to(dev)
    accum, denom = torch.rand(P, 1, generator=g).to(dev), torch.randint(0, 5, (P, 1), generator=g).float().to(dev)
    maxr = torch.randint(0, 30, (P,), generator=g).float().to(dev)
    a2, d2, m2 = accum.clone(), denom.clone(), maxr.clone()
    vis = radii > 0                                               # reference lines, verbatim semantics
    m2[vis] = torch.max(m2[vis], radii[vis])
    a2[vis] += torch.norm(grad[vis], dim=-1, keepdim=True)
    d2[vis] += 1
    densification_stats(accum, denom, maxr, grad, radii)
    assert torch.equal(maxr, m2) and torch.equal(denom, d2)
    torch.testing.assert_close(accum, a2, rtol=1e-6, atol=1e-7)
