"""f4 on the device: PLY rows -> rasterizer inputs and back, against a restatement of the reference's
load_ply / save_ply lines and getters (/root/reference/scene/gaussian_model.py:35-41, :95-115, :192-255)."""
import numpy as np
import pytest
import torch

import surfel_ply as PLY
from test_ply_cpu import random_model, reference_file_bytes

gpu = pytest.mark.gpu


def reference_load(blob):
    """load_ply restated: columns by name -> parameter tensors (float32)."""
    count, names, offset = PLY.parse_header(blob)
    data = np.frombuffer(blob, dtype="<f4", offset=offset).reshape(count, len(names))
    col = {n: data[:, i] for i, n in enumerate(names)}
    xyz = np.stack((col["x"], col["y"], col["z"]), axis=1)
    dc = np.zeros((count, 3, 1), "f4")
    for c in range(3):
        dc[:, c, 0] = col[f"f_dc_{c}"]
    extra = np.stack([col[f"f_rest_{i}"] for i in range(45)], axis=1).reshape(count, 3, 15)
    t = lambda a: torch.tensor(a, dtype=torch.float)
    return dict(xyz=t(xyz), features_dc=t(dc).transpose(1, 2).contiguous(), features_rest=t(extra).transpose(1, 2).contiguous(),
                opacity=t(col["opacity"][:, None]), scaling=t(np.stack((col["scale_0"], col["scale_1"]), 1)),
                rotation=t(np.stack([col[f"rot_{i}"] for i in range(4)], 1)))


@gpu
@pytest.mark.parametrize("P", [1, 31, 1000, 40001])
@pytest.mark.parametrize("shuffle", [False, True])
def test_load_ply_matches_reference(tmp_path, P, shuffle):
    model = random_model(P, 3 + P)
    names = None
    if shuffle:
        names = list(np.random.default_rng(1).permutation(PLY.reference_attributes())) + []
        names = [str(n) for n in names]
    blob = reference_file_bytes(*model, names=names)
    path = tmp_path / "point_cloud.ply"
    path.write_bytes(blob)
    ref = reference_load(blob)
    raw = PLY.load_ply(str(path), activate=False)
    assert torch.equal(raw["means3D"].cpu(), ref["xyz"])
    assert torch.equal(raw["shs"].cpu(), torch.cat((ref["features_dc"], ref["features_rest"]), dim=1))    # get_features
    assert torch.equal(raw["opacities"].cpu(), ref["opacity"])
    assert torch.equal(raw["scales"].cpu(), ref["scaling"]) and torch.equal(raw["rotations"].cpu(), ref["rotation"])
    act = PLY.load_ply(str(path), activate=True)
    assert torch.equal(act["means3D"], raw["means3D"]) and torch.equal(act["shs"], raw["shs"])
    torch.testing.assert_close(act["opacities"].cpu(), torch.sigmoid(ref["opacity"]), rtol=2e-6, atol=1e-7)
    torch.testing.assert_close(act["scales"].cpu(), torch.exp(ref["scaling"]), rtol=2e-6, atol=0)
    torch.testing.assert_close(act["rotations"].cpu(), torch.nn.functional.normalize(ref["rotation"]), rtol=2e-6, atol=1e-7)


@gpu
@pytest.mark.parametrize("P", [1, 33, 5000])
def test_save_ply_is_byte_identical_and_round_trips(tmp_path, P):
    model = random_model(P, 11 + P)
    dev = torch.device("cuda")
    tensors = [torch.from_numpy(a).to(dev) for a in model]
    path = tmp_path / "out" / "point_cloud.ply"
    PLY.save_ply(str(path), *tensors)
    assert path.read_bytes() == reference_file_bytes(*model)
    back = PLY.load_ply(str(path), activate=False)
    assert torch.equal(back["means3D"], tensors[0])
    assert torch.equal(back["shs"], torch.cat((tensors[1], tensors[2]), dim=1))
    assert torch.equal(back["opacities"], tensors[3]) and torch.equal(back["scales"], tensors[4])
    assert torch.equal(back["rotations"], tensors[5])


@gpu
def test_loaded_model_renders(tmp_path):
    """A saved model goes through the op unchanged: file -> load_ply(activate=True) -> GaussianRasterizer."""
    import surfel_scenes as S
    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    dev = torch.device("cuda")
    scene, cam = S.named("config1")
    P, W, H = S.CONFIGS["config1"]
    inv_sig = lambda y: torch.log(y / (1 - y))
    params = [scene["means3D"], scene["shs"][:, :1].contiguous(), scene["shs"][:, 1:].contiguous(),
              inv_sig(scene["opacities"].clamp(1e-4, 1 - 1e-4)), torch.log(scene["scales"]), scene["rotations"]]
    path = tmp_path / "scene.ply"
    PLY.save_ply(str(path), *[t.to(dev) for t in params])
    m = PLY.load_ply(str(path), activate=True)
    rs = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"].to(dev), projmatrix=cam["projmatrix"].to(dev), sh_degree=3,
        campos=cam["campos"].to(dev), prefiltered=False, debug=False)
    rast = GaussianRasterizer(rs)
    with torch.no_grad():
        a = rast(means3D=m["means3D"], means2D=torch.zeros(P, 3, device=dev), shs=m["shs"], opacities=m["opacities"],
                 scales=m["scales"], rotations=m["rotations"])
        b = rast(means3D=scene["means3D"].to(dev), means2D=torch.zeros(P, 3, device=dev), shs=scene["shs"].to(dev),
                 opacities=scene["opacities"].to(dev), scales=scene["scales"].to(dev),
                 rotations=torch.nn.functional.normalize(scene["rotations"]).to(dev))
    assert float(a[2][1].max()) > 0.0
    torch.testing.assert_close(a[0], b[0], rtol=0, atol=2e-4)
