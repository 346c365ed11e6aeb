"""Property tests (hypothesis) of the host-side logic either side of the path: band / view sharding and
the PLY header + column table.  No GPU."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

import surfel_parallel as SP
import surfel_ply as PLY


@settings(max_examples=200, deadline=None)
@given(H=st.integers(1, 20000), world=st.integers(1, 64))
def test_tile_row_bands_partition_the_frame(H, world):
    gy = SP.tile_rows(H)
    bands = [SP.tile_row_band(H, r, world) for r in range(world)]
    assert bands[0][0] == 0 and bands[-1][1] == gy
    assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
    sizes = [e - b for b, e in bands]
    assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1 and sum(sizes) == gy
    # pixel rows of the bands tile [0, H) as well
    rows = [SP.band_pixel_rows(H, b) for b in bands]
    assert rows[0][0] == 0 and rows[-1][1] == H and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 500), world=st.integers(1, 64))
def test_views_are_dealt_exactly_once(n, world):
    dealt = sorted(v for r in range(world) for v in SP.shard_views(n, r, world))
    assert dealt == list(range(n))
    assert max(len(SP.shard_views(n, r, world)) for r in range(world)) - min(len(SP.shard_views(n, r, world)) for r in range(world)) <= 1


@settings(max_examples=100, deadline=None)
@given(count=st.integers(0, 10**7), perm=st.permutations(PLY.reference_attributes()),
       extra=st.lists(st.sampled_from(["confidence", "label", "f_extra_7"]), unique=True, max_size=3),
       comment=st.booleans())
def test_header_round_trip_and_column_table_under_any_property_order(count, perm, extra, comment):
    names = list(perm) + list(extra)                       # files may carry properties the op does not use
    head = PLY.header_bytes(count, names)
    if comment:
        head = head.replace(b"ply\n", b"ply\ncomment written by something else\n", 1)
    got_count, got_names, offset = PLY.parse_header(head + b"\x00" * 16)
    assert (got_count, got_names, offset) == (count, names, len(head))
    cols = PLY.column_table(names)
    ref = PLY.reference_attributes()
    ref_cols = PLY.column_table(ref)
    assert [names[c] for c in cols] == [ref[c] for c in ref_cols]          # same properties, wherever they sit
    assert len(set(cols)) == 58 and all(names[c] not in ("nx", "ny", "nz") + tuple(extra) for c in cols)


@settings(max_examples=50, deadline=None)
@given(P=st.integers(1, 40), seed=st.integers(0, 2**31 - 1))
def test_rows_gathered_through_the_table_are_the_saved_parameters(P, seed):
    """numpy twin of surfel_ply_unpack(activate=0) on rows assembled like save_ply: the table must return the
    parameters that went in, in the (P,16,3) coefficient-major SH layout."""
    from test_ply_cpu import random_model, reference_file_bytes
    xyz, dc, rest, opa, scale, rot = random_model(P, seed)
    blob = reference_file_bytes(xyz, dc, rest, opa, scale, rot)
    count, names, offset = PLY.parse_header(blob)
    rows = np.frombuffer(blob, dtype="<f4", offset=offset).reshape(count, len(names))
    t = rows[:, PLY.column_table(names)]
    assert np.array_equal(t[:, :3], xyz) and np.array_equal(t[:, 3:51].reshape(P, 16, 3), np.concatenate((dc, rest), 1))
    assert np.array_equal(t[:, 51:52], opa) and np.array_equal(t[:, 52:54], scale) and np.array_equal(t[:, 54:58], rot)
