"""The PyTorch restatements that the f1 / f2 GPU parity tests compare the CUDA kernels with
(tests/test_postprocess_gpu.py::reference_tail, tests/test_loss_gpu.py::reference_loss) are themselves
pinned here to THE REFERENCE'S OWN PYTHON: tests/golden/ref_tail_loss.npz was produced by running the
reference's unmodified render() tail and loss_utils (tests/golden/make_golden_tail_loss.py).  Chain of
evidence: reference == golden (generated in the build container), golden == restatement (this file, CPU),
restatement == CUDA (the -m gpu tests)."""
import os
import types

import numpy as np
import pytest
import torch

from test_loss_gpu import reference_loss
from test_postprocess_gpu import reference_tail

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_tail_loss.npz")
KEYS = ("rend_alpha", "rend_normal", "rend_dist", "surf_depth", "surf_normal")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("ratio", [0.0, 1.0, 0.3])
def test_tail_restatement_equals_reference_render_tail(gold, ratio):
    t = lambda k: torch.from_numpy(gold[k])
    cam = types.SimpleNamespace(world_view_transform=t("viewmatrix"), full_proj_transform=t("projmatrix"),
                                image_width=int(gold["W"]), image_height=int(gold["H"]))
    allmap = t("allmap").clone().requires_grad_(True)
    out = reference_tail(allmap, cam, ratio)
    tag = str(ratio).replace(".", "p")
    for k in KEYS:
        torch.testing.assert_close(out[k].detach(), t(f"tail_{tag}_{k}"), rtol=1e-5, atol=2e-6, msg=lambda m, k=k: f"{k}: {m}")
    sum((out[k] * t("cot_" + k)).sum() for k in KEYS).backward()
    ref = t(f"tail_{tag}_grad_allmap")
    # where alpha == 0 the reference's own gradient of D/alpha is NaN (0 * inf behind nan_to_num): same pixels, same NaNs
    assert torch.equal(torch.isnan(allmap.grad), torch.isnan(ref)) and bool(torch.isnan(ref).any())
    scale = float(torch.nan_to_num(ref, 0.0, 0.0, 0.0).abs().max())
    torch.testing.assert_close(allmap.grad, ref, rtol=1e-4, atol=1e-5 * scale, equal_nan=True)


@pytest.mark.parametrize("lam", [0.2, 1.0, 0.0])
def test_loss_restatement_equals_reference_loss_utils(gold, lam):
    img = torch.from_numpy(gold["loss_img"]).clone().requires_grad_(True)
    gt = torch.from_numpy(gold["loss_gt"])
    loss = reference_loss(img, gt, lam)
    loss.backward()
    tag = str(lam).replace(".", "p")
    assert abs(float(loss) - float(gold[f"loss_{tag}_value"])) < 1e-6
    ref = torch.from_numpy(gold[f"loss_{tag}_grad"])
    torch.testing.assert_close(img.grad, ref, rtol=1e-4, atol=1e-6 * max(1.0, float(ref.abs().max())))


def test_l1_and_ssim_values(gold):
    img, gt = torch.from_numpy(gold["loss_img"]), torch.from_numpy(gold["loss_gt"])
    assert abs(float(reference_loss(img, gt, 0.0)) - float(gold["l1_value"])) < 1e-7
    assert abs((1.0 - float(reference_loss(img, gt, 1.0))) - float(gold["ssim_value"])) < 1e-6
