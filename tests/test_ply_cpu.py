"""f4 host logic: PLY header text, parsing and the property-name -> column table (no GPU).

Oracle: a numpy restatement of the reference's save_ply (/root/reference/scene/gaussian_model.py:192-209):
a structured array with one 'f4' field per attribute, which plyfile writes as
`property float <name>` lines and raw little-endian rows."""
import numpy as np
import pytest

import surfel_ply as PLY


def reference_file_bytes(xyz, f_dc, f_rest, opacity, scale, rot, names=None):
    """What save_ply puts on disk (features given as (P,1,3) and (P,15,3))."""
    P = xyz.shape[0]
    dc = np.transpose(f_dc, (0, 2, 1)).reshape(P, -1)             # transpose(1, 2).flatten(start_dim=1)
    rest = np.transpose(f_rest, (0, 2, 1)).reshape(P, -1)
    attributes = np.concatenate((xyz, np.zeros_like(xyz), dc, rest, opacity, scale, rot), axis=1).astype("<f4")
    ref_names = PLY.reference_attributes()
    if names is not None:                                           # same data, another property order
        attributes = attributes[:, [ref_names.index(n) for n in names]]
    head = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % P
    head += "".join("property float %s\n" % n for n in (names or ref_names)) + "end_header\n"
    return head.encode("ascii") + np.ascontiguousarray(attributes).tobytes()


def random_model(P, seed):
    r = np.random.default_rng(seed)
    return (r.normal(size=(P, 3)).astype("f4"), r.normal(size=(P, 1, 3)).astype("f4"), r.normal(size=(P, 15, 3)).astype("f4"),
            r.normal(size=(P, 1)).astype("f4"), r.normal(size=(P, 2)).astype("f4"), r.normal(size=(P, 4)).astype("f4"))


def test_attribute_list_matches_reference_order():
    names = PLY.reference_attributes()
    assert len(names) == PLY.ROW_FLOATS == 61
    assert names[:9] == ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"]
    assert names[9] == "f_rest_0" and names[53] == "f_rest_44"
    assert names[54:] == ["opacity", "scale_0", "scale_1", "rot_0", "rot_1", "rot_2", "rot_3"]


def test_header_round_trip():
    blob = reference_file_bytes(*random_model(5, 0))
    assert blob.startswith(PLY.header_bytes(5))
    count, names, offset = PLY.parse_header(blob)
    assert count == 5 and names == PLY.reference_attributes() and offset == len(PLY.header_bytes(5))
    assert len(blob) == offset + 5 * 61 * 4
    # comments and CRLF line ends, as other writers produce them
    alt = blob[:4] + b"comment made elsewhere\r\n" + blob[4:offset].replace(b"\n", b"\r\n") + blob[offset:]
    c2, n2, o2 = PLY.parse_header(alt)
    assert (c2, n2) == (count, names) and alt[o2:] == blob[offset:]


def test_column_table_follows_names_not_positions():
    names = PLY.reference_attributes()
    cols = PLY.column_table(names)
    assert len(cols) == 58 and cols[:3] == [0, 1, 2]
    # shs[k][c]: k = 0 -> f_dc_c (columns 6..8); k >= 1 -> f_rest_{c*15 + k-1} (column 9 + ...)
    assert cols[3:6] == [6, 7, 8]
    for k in range(1, 16):
        for c in range(3):
            assert cols[3 + 3 * k + c] == 9 + c * 15 + (k - 1)
    assert cols[51:] == [54, 55, 56, 57, 58, 59, 60]
    shuffled = list(reversed(names))
    cols2 = PLY.column_table(shuffled)
    assert [shuffled[c] for c in cols2] == [names[c] for c in cols]


@pytest.mark.parametrize("mutate, message", [
    (lambda b: b.replace(b"binary_little_endian", b"ascii"), "unsupported PLY format"),
    (lambda b: b.replace(b"property float opacity", b"property double opacity"), "only float32"),
    (lambda b: b.replace(b"element vertex", b"element face"), "expected 'vertex'"),
    (lambda b: b[4:], "missing magic"),
    (lambda b: b.replace(b"end_header", b"end_of_head"), "no end_header"),
])
def test_rejected_headers(mutate, message):
    blob = reference_file_bytes(*random_model(2, 1))
    with pytest.raises(ValueError, match=message):
        PLY.parse_header(mutate(blob))


def test_missing_or_wrong_properties():
    names = PLY.reference_attributes()
    with pytest.raises(ValueError, match="f_rest"):
        PLY.column_table([n for n in names if n != "f_rest_44"])            # a degree-2 file
    with pytest.raises(ValueError, match="opacity"):
        PLY.column_table([n for n in names if n != "opacity"])
    with pytest.raises(ValueError, match="scale"):
        PLY.column_table(names + ["scale_2"])                                # a 3DGS (3-scale) model


def test_no_cpu_path(tmp_path):
    import torch
    path = tmp_path / "m.ply"
    path.write_bytes(reference_file_bytes(*random_model(3, 2)))
    with pytest.raises(RuntimeError, match="no CPU path"):
        PLY.load_ply(str(path), device="cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        PLY.unpack_rows(torch.zeros(3, 61), PLY.reference_attributes(), True)


# ---- pinned to the reference's own save_ply / load_ply (tests/golden/make_golden_ply.py) ----

def _gold():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_ply.npz"))


def test_attribute_order_is_the_references():
    g = _gold()
    assert [str(n) for n in g["names"]] == PLY.reference_attributes()


def test_row_assembly_matches_reference_save_ply():
    """reference_file_bytes (the oracle of the byte-identity GPU test) assembles rows exactly like save_ply."""
    g = _gold()
    blob = reference_file_bytes(g["xyz"], g["features_dc"], g["features_rest"], g["opacity"], g["scaling"], g["rotation"])
    count, names, offset = PLY.parse_header(blob)
    rows = np.frombuffer(blob, dtype="<f4", offset=offset).reshape(count, len(names))
    assert count == g["rows"].shape[0] and np.array_equal(rows, g["rows"])


def test_column_table_reproduces_reference_load_ply():
    """Gathering the golden rows through column_table gives the tensors load_ply builds (features in the
    (P,16,3) layout = cat(_features_dc, _features_rest)), and the activations are the getters'."""
    import torch
    g = _gold()
    cols = PLY.column_table([str(n) for n in g["names"]])
    t = g["rows"][:, cols]                                     # (P, 58) in surfel_ply_unpack's target order
    assert np.array_equal(t[:, 0:3], g["loaded_xyz"])
    shs = t[:, 3:51].reshape(-1, 16, 3)
    assert np.array_equal(shs[:, :1], g["loaded_features_dc"]) and np.array_equal(shs[:, 1:], g["loaded_features_rest"])
    assert np.array_equal(shs, g["act_features"])
    assert np.array_equal(t[:, 51:52], g["loaded_opacity"]) and np.array_equal(t[:, 52:54], g["loaded_scaling"])
    assert np.array_equal(t[:, 54:58], g["loaded_rotation"])
    # what activate=1 must produce (gaussian_model.py getters), here with torch on CPU
    np.testing.assert_allclose(torch.sigmoid(torch.from_numpy(t[:, 51:52])).numpy(), g["act_opacity"], rtol=1e-6)
    np.testing.assert_allclose(np.exp(t[:, 52:54]), g["act_scaling"], rtol=1e-6)
    np.testing.assert_allclose(torch.nn.functional.normalize(torch.from_numpy(t[:, 54:58])).numpy(), g["act_rotation"], rtol=1e-6)
