"""Pins the f4 (model PLY) restatements to THE REFERENCE'S OWN save_ply / load_ply.

`plyfile` is not installed here, so the byte-level container cannot be produced by the reference; what CAN
be pinned is everything the reference itself decides: the attribute list and order, the row assembly of
save_ply (/root/reference/scene/gaussian_model.py:176-209: transposes, flattening, zero normals) and the
column -> tensor mapping of load_ply (:215-255).  A stub `plyfile` captures the structured array save_ply
hands to PlyElement.describe and serves it back to load_ply.  Writes tests/golden/ref_ply.npz.

Usage:  python tests/golden/make_golden_ply.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "2d-gaussian-splatting_b200"))
REF = "/root/reference"


def main():
    import make_golden as MG
    MG.cpu_patches()
    MG.stub_modules({})
    store = {}

    class Prop:
        def __init__(self, name):
            self.name = name

    class Element:
        def __init__(self, data):
            self.data = data
            self.properties = [Prop(n) for n in data.dtype.names]

        def __getitem__(self, name):
            return self.data[name]

    class PlyElement:
        @staticmethod
        def describe(elements, name):
            assert name == "vertex"
            store["elements"] = elements.copy()
            return Element(elements)

    class PlyData:
        def __init__(self, elements=None):
            self.elements = elements or []

        def write(self, path):
            store["written_to"] = path

        @staticmethod
        def read(path):
            return PlyData([Element(store["elements"])])

    ply = types.ModuleType("plyfile")
    ply.PlyData, ply.PlyElement = PlyData, PlyElement
    sys.modules["plyfile"] = ply
    sys.path.insert(0, REF)
    import scene.gaussian_model as GM
    GM.PlyData, GM.PlyElement = PlyData, PlyElement       # the module did `from plyfile import PlyData, PlyElement`
    GM.mkdir_p = lambda p: None

    P = 7
    g = torch.Generator("cpu").manual_seed(99)
    pc = GM.GaussianModel(3)
    pc._xyz = torch.randn(P, 3, generator=g)
    pc._features_dc = torch.randn(P, 1, 3, generator=g)
    pc._features_rest = torch.randn(P, 15, 3, generator=g)
    pc._opacity = torch.randn(P, 1, generator=g)
    pc._scaling = torch.randn(P, 2, generator=g)
    pc._rotation = torch.randn(P, 4, generator=g)
    names = pc.construct_list_of_attributes()
    pc.save_ply("/nonexistent/point_cloud.ply")
    el = store["elements"]
    assert list(el.dtype.names) == names and all(el.dtype[n] == np.dtype("f4") for n in names)
    rows = np.stack([el[n] for n in names], axis=1).astype(np.float32)

    back = GM.GaussianModel(3)
    back.load_ply("/nonexistent/point_cloud.ply")
    np.savez_compressed(
        os.path.join(HERE, "ref_ply.npz"), names=np.array(names), rows=rows,
        xyz=pc._xyz.numpy(), features_dc=pc._features_dc.numpy(), features_rest=pc._features_rest.numpy(),
        opacity=pc._opacity.numpy(), scaling=pc._scaling.numpy(), rotation=pc._rotation.numpy(),
        loaded_xyz=back._xyz.detach().numpy(), loaded_features_dc=back._features_dc.detach().numpy(),
        loaded_features_rest=back._features_rest.detach().numpy(), loaded_opacity=back._opacity.detach().numpy(),
        loaded_scaling=back._scaling.detach().numpy(), loaded_rotation=back._rotation.detach().numpy(),
        act_opacity=back.get_opacity.detach().numpy(), act_scaling=back.get_scaling.detach().numpy(),
        act_rotation=back.get_rotation.detach().numpy(), act_features=back.get_features.detach().numpy())
    print("wrote ref_ply.npz:", len(names), "attributes,", rows.shape)


if __name__ == "__main__":
    main()
