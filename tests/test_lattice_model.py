"""The lattice formulation of the vertex weld (tests/lattice_model.py) equals the reference's
sort-and-compact (oracle/orc_marching.c) bit for bit: same vertices in the same order, same
external keys, same triangles.  This is the algorithmic identity the CUDA marching kernels rely on."""
import numpy as np
import pytest

from lattice_model import lattice_mesh
from oracle.oracle import Oracle
from test_oracle_marching import MARCHING_CASES, sphere_field, to_image


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def check(orc, vol, z_first, z_last, key_offset=(0, 0, 0)):
    d, h, w = vol.shape
    img, zs = to_image(vol)
    ref = orc.marching_generate(img, w, h, zs, 0, z_first, z_last, key_offset)
    got = lattice_mesh(vol, orc.marching_tables(), z_first, z_last, key_offset)
    assert got["num_internal"] == ref["num_internal"]
    np.testing.assert_array_equal(got["vertices"].view(np.uint32), ref["vertices"].view(np.uint32))
    ni = ref["num_internal"]
    np.testing.assert_array_equal(got["keys"][ni:], ref["keys"][ni:])
    np.testing.assert_array_equal(got["external"], np.arange(ni, len(ref["vertices"])))
    np.testing.assert_array_equal(got["triangles"], ref["triangles"])
    return ref


@pytest.mark.parametrize("name", sorted(MARCHING_CASES))
def test_reference_volumes(orc, name):
    vol = MARCHING_CASES[name]()
    if name != "alternating":
        vol = vol[:24, :40, :37].copy()      # the python triangle loop is slow; keep the cases small
    d = vol.shape[0]
    ref = check(orc, vol, 0, d - 1)
    assert len(ref["vertices"]) > ref["num_internal"] > 0
    # shipped out in two blocks: the seam plane is external in both
    check(orc, vol, 0, d // 2)
    check(orc, vol, d // 2, d - 1)


def test_nan_holes_offsets_and_random_fields(orc):
    vol = sphere_field(14, 13, 12, 6.5, 6.0, 5.5, 7.2)          # truncated by the box on every side
    check(orc, vol, 0, 11, (7, 100, 3))
    vol[5, 2, 5] = np.nan
    vol[0, 0, 0] = np.inf
    vol[11, 12, 13] = np.nan
    check(orc, vol, 0, 11)
    check(orc, vol, 3, 9, (1, 2, 3))
    rng = np.random.RandomState(7)
    for _ in range(4):
        vol = rng.standard_normal((7, 9, 35)).astype(np.float32)   # 34 cells: spans two 32-cell words
        vol[rng.rand(*vol.shape) < 0.15] = np.nan
        vol[rng.rand(*vol.shape) < 0.05] = 0.0                     # exact zeros count as outside (>= 0)
        check(orc, vol, 0, 6)
        check(orc, vol, 2, 5, (3, 0, 9))
