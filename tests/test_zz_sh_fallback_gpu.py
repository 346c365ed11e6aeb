"""preprocess_fwd's non-prefetched SH path (runs last: file name sorts after the other GPU tests).

The kernel requests SH rows by cp.async only for splats whose centre projects within 1.5x the screen; a
large splat centred further out that still touches tiles gets its row through the plain fallback.
tests/offscreen_scene.py builds such splats (tests/test_oracle_cpu.py checks on the CPU that the scene really
contains >= 30 of them); what preprocess produces must still be bit-exact (radii, tile counts, R, clamp bits) and
within 1e-6 (RGB from SH) of the oracle, for those rows and for the prefetched rows sharing their warps."""
import numpy as np
import pytest

from offscreen_scene import centre_outside_margin, offscreen_scene
from test_parity_gpu import run_both

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sh_degree", [3, 0])
def test_sh_rows_of_offscreen_centred_splats(oracle, cuda_lib, sh_degree):
    scene, cam = offscreen_scene()
    bg = np.array([0.05, 0.1, 0.2], np.float32)
    pre, binned, img, pipe = run_both(oracle, scene, cam, bg, sh_degree)
    got = pipe.preprocess()
    vis = pre["radii"] > 0
    far = centre_outside_margin(scene, cam)
    assert int((vis & far).sum()) >= 30
    np.testing.assert_array_equal(got["radii"], pre["radii"])
    np.testing.assert_array_equal(got["tiles_touched"], pre["tiles_touched"])
    assert got["R"] == binned["R"]
    np.testing.assert_allclose(got["rgb"][vis & far], pre["rgb"][vis & far], atol=1e-6, rtol=0)     # the fallback rows
    np.testing.assert_allclose(got["rgb"][vis & ~far], pre["rgb"][vis & ~far], atol=1e-6, rtol=0)   # the prefetched rows
    np.testing.assert_array_equal(got["clamped"][vis], pre["clamped"][vis])
