"""Trained-model PLY files of the reference, read and written for the rasterizer.

The reference saves a model with `GaussianModel.save_ply` and restores it with `load_ply`
(/root/reference/scene/gaussian_model.py:176-209, :215-255): a binary little-endian PLY, element
`vertex`, 61 float32 properties per splat — x y z nx ny nz f_dc_0..2 f_rest_0..44 opacity scale_0..1
rot_0..3 — all PRE-activation, SH coefficients channel-major (f_rest_{c*15+k}).

Host side (this file): header text, property-name -> column table, file IO.  Device side
(csrc/ply_pack.cu through the C ABI): ONE kernel turns the raw rows into the op's inputs
(`surfel_ply_unpack`: transposes SH to the (P,16,3) coefficient-major layout, applies sigmoid / exp /
normalise when asked) and one gathers parameters back into rows (`surfel_ply_pack`).  The reference does
this with ~70 numpy column copies on the CPU plus six host->device tensor constructions.

There is no CPU path for the row shuffling: tensors live on a CUDA device.
"""
import ctypes
import os

import numpy as np
import torch

ROW_FLOATS = 61
SH_COEFFS = 16          # degree 3; the rasterizer's vectorised layout


def reference_attributes():
    """Property names in the order save_ply writes them (gaussian_model.py:176-190)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(3)]
    names += [f"f_rest_{i}" for i in range(3 * (SH_COEFFS - 1))]
    names += ["opacity"] + [f"scale_{i}" for i in range(2)] + [f"rot_{i}" for i in range(4)]
    return names


def header_bytes(count, names=None):
    """The header plyfile writes for a structured array of `f4` fields (what save_ply produces)."""
    names = reference_attributes() if names is None else names
    lines = ["ply", "format binary_little_endian 1.0", f"element vertex {int(count)}"]
    lines += [f"property float {n}" for n in names]
    lines.append("end_header")
    return ("\n".join(lines) + "\n").encode("ascii")


def parse_header(buf):
    """bytes (at least the header) -> (vertex count, property names in file order, offset of the data).

    Accepts what the reference's loader accepts for this format: the `vertex` element first, every
    property a 4-byte float.  Raises ValueError otherwise (ascii / big-endian files, list properties,
    doubles): such files never come out of save_ply."""
    end = buf.find(b"end_header")
    if end < 0:
        raise ValueError("not a PLY file: no end_header")
    nl = buf.find(b"\n", end)
    if nl < 0:
        raise ValueError("truncated PLY header")
    lines = buf[:end].decode("ascii", errors="replace").replace("\r\n", "\n").split("\n")
    if not lines or lines[0].strip() != "ply":
        raise ValueError("not a PLY file: missing magic")
    fmt, count, names, element = None, None, [], None
    for line in lines[1:]:
        tok = line.split()
        if not tok or tok[0] in ("comment", "obj_info"):
            continue
        if tok[0] == "format":
            fmt = tok[1:]
        elif tok[0] == "element":
            if element is None and tok[1] != "vertex":
                raise ValueError(f"first element is '{tok[1]}', expected 'vertex'")
            element = tok[1]
            if element == "vertex":
                count = int(tok[2])
        elif tok[0] == "property" and element == "vertex":
            if tok[1] not in ("float", "float32"):
                raise ValueError(f"property '{tok[-1]}' has type '{tok[1]}'; only float32 properties are supported")
            names.append(tok[2])
    if fmt != ["binary_little_endian", "1.0"]:
        raise ValueError(f"unsupported PLY format {fmt}; save_ply writes binary_little_endian 1.0")
    if count is None:
        raise ValueError("PLY header has no vertex element")
    return count, names, nl + 1


def column_table(names):
    """Column of every target float of `surfel_ply_unpack` (58 entries; include/surfel_rasterizer.h),
    addressed by property NAME like load_ply (gaussian_model.py:215-247)."""
    pos = {n: i for i, n in enumerate(names)}
    rest = sorted((n for n in names if n.startswith("f_rest_")), key=lambda n: int(n.split("_")[-1]))
    if len(rest) != 3 * (SH_COEFFS - 1):       # the reference asserts the same for its max_sh_degree (3)
        raise ValueError(f"{len(rest)} f_rest_* properties; a degree-3 model has {3 * (SH_COEFFS - 1)}")
    scale = sorted((n for n in names if n.startswith("scale_")), key=lambda n: int(n.split("_")[-1]))
    rot = sorted((n for n in names if n.startswith("rot")), key=lambda n: int(n.split("_")[-1]))
    if len(scale) != 2 or len(rot) != 4:
        raise ValueError(f"expected 2 scale_* and 4 rot_* properties, found {len(scale)} and {len(rot)}")
    try:
        cols = [pos["x"], pos["y"], pos["z"]]
        for k in range(SH_COEFFS):
            for c in range(3):
                cols.append(pos[f"f_dc_{c}"] if k == 0 else pos[rest[c * (SH_COEFFS - 1) + (k - 1)]])
        cols.append(pos["opacity"])
    except KeyError as e:
        raise ValueError(f"PLY file lacks property {e}") from None
    cols += [pos[n] for n in scale] + [pos[n] for n in rot]
    return cols


def _cabi():
    from diff_surfel_rasterization import _cabi
    return _cabi


def unpack_rows(rows, names, activate):
    """rows: (P, len(names)) float32 CUDA tensor of raw PLY rows -> dict of tensors (see load_ply)."""
    if not rows.is_cuda or rows.dtype != torch.float32 or not rows.is_contiguous():
        raise RuntimeError("unpack_rows: rows must be a contiguous float32 CUDA tensor (no CPU path)")
    cabi = _cabi()
    lib = cabi.load()
    P, row_floats = rows.shape
    dev = rows.device
    cols = (ctypes.c_int32 * 58)(*column_table(names))
    out = {"means3D": torch.empty(P, 3, device=dev), "shs": torch.empty(P, SH_COEFFS, 3, device=dev),
           "opacities": torch.empty(P, 1, device=dev), "scales": torch.empty(P, 2, device=dev),
           "rotations": torch.empty(P, 4, device=dev)}
    with torch.cuda.device(dev):
        cabi.check(lib.surfel_ply_unpack(P, row_floats, rows.data_ptr(), cols, int(bool(activate)),
                                         out["means3D"].data_ptr(), out["shs"].data_ptr(), out["opacities"].data_ptr(),
                                         out["scales"].data_ptr(), out["rotations"].data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream))
    return out


def load_ply(path, device="cuda", activate=True):
    """Read a model saved by the reference.

    activate=True  -> the rasterizer's inputs: means3D (P,3), shs (P,16,3), opacities (P,1) in (0,1),
                      scales (P,2) > 0, rotations (P,4) unit (what GaussianModel's getters return).
    activate=False -> the stored parameters in the same layout (shs[:, :1] / shs[:, 1:] are
                      _features_dc / _features_rest of load_ply)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("load_ply: a CUDA device is required (no CPU path)")
    with open(path, "rb") as f:
        head = f.read(1 << 16)
    count, names, offset = parse_header(head)
    need = offset + count * len(names) * 4
    if os.path.getsize(path) < need:
        raise ValueError(f"{path}: {os.path.getsize(path)} bytes, header promises {need}")
    host = np.fromfile(path, dtype="<f4", count=count * len(names), offset=offset).reshape(count, len(names))
    rows = torch.from_numpy(host).pin_memory().to(device, non_blocking=True)
    return unpack_rows(rows, names, activate)


def pack_rows(xyz, features_dc, features_rest, opacity, scaling, rotation):
    """Parameters (as GaussianModel holds them) -> (P,61) float32 CUDA rows in save_ply's column order."""
    cabi = _cabi()
    lib = cabi.load()
    ts = [t.detach().contiguous().float() for t in (xyz, features_dc, features_rest, opacity, scaling, rotation)]
    if not all(t.is_cuda for t in ts):
        raise RuntimeError("pack_rows: CUDA tensors required (no CPU path)")
    P = ts[0].shape[0]
    want = [(P, 3), (P, 1, 3), (P, SH_COEFFS - 1, 3), (P, 1), (P, 2), (P, 4)]
    for t, w in zip(ts, want):
        if tuple(t.shape) != w:
            raise RuntimeError(f"pack_rows: tensor of shape {tuple(t.shape)}, expected {w}")
    dev = ts[0].device
    rows = torch.empty(P, ROW_FLOATS, device=dev)
    with torch.cuda.device(dev):
        cabi.check(lib.surfel_ply_pack(P, *[t.data_ptr() for t in ts], rows.data_ptr(),
                                       torch.cuda.current_stream(dev).cuda_stream))
    return rows


def save_ply(path, xyz, features_dc, features_rest, opacity, scaling, rotation):
    """Write the file GaussianModel.save_ply would write for these parameters (byte for byte)."""
    rows = pack_rows(xyz, features_dc, features_rest, opacity, scaling, rotation).cpu().numpy()
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "wb") as f:
        f.write(header_bytes(rows.shape[0]))
        f.write(rows.astype("<f4", copy=False).tobytes())
