"""GPU marching tetrahedra + weld vs the oracle: vertices, keys, triangles identical; the
reference's end-to-end cases (test/test_marching.cpp:594-632) give manifolds."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.oracle import Oracle
from fixtures import is_manifold
from gpu_common import assert_block_equal
from test_oracle_marching import MARCHING_CASES, to_image, weld_by_key, sphere_field


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def oracle_mesh(orc, vol, key_offset=(0, 0, 0), z0=0, z1=None):
    d, h, w = vol.shape
    img, zs = to_image(vol)
    return orc.marching_generate(img, w, h, zs, 0, z0, d - 1 if z1 is None else z1, key_offset)


@pytest.mark.parametrize("name", sorted(MARCHING_CASES))
def test_reference_cases(orc, name):
    import mlsgpu_b200 as m
    vol = MARCHING_CASES[name]()
    ref = oracle_mesh(orc, vol)
    with m.Worker(max_splats=16, max_cells=82) as w:
        blocks = w.marching_generate_host_field(vol, (0, 0, 0))
    assert len(blocks) == 1
    assert_block_equal(blocks[0], ref)
    assert is_manifold(len(blocks[0]["vertices"]), blocks[0]["triangles"]) == ""


def test_key_offset_and_scale_bias(orc):
    import mlsgpu_b200 as m
    vol = sphere_field(40, 36, 33, 19.5, 17.0, 16.0, 21.0)
    ko = (7, 100, 3)
    ref = oracle_mesh(orc, vol, ko)
    with m.Worker(max_splats=16, max_cells=63) as w:
        w.set_scale_bias(0.0125, (10.0, -20.0, 30.0))
        blocks = w.marching_generate_host_field(vol, ko)
    assert len(blocks) == 1 and len(ref["vertices"]) > ref["num_internal"] > 0
    assert_block_equal(blocks[0], ref, 0.0125, (10.0, -20.0, 30.0))


def test_minimum_mesh_memory_splits_and_welds(orc):
    """Minimum legal mesh memory (test/test_marching.cpp:561-565): the bin is shipped out in
    several blocks whose seams weld by key into the same manifold."""
    import mlsgpu_b200 as m
    vol = MARCHING_CASES["truncated"]()
    d, h, w_ = vol.shape
    ref = oracle_mesh(orc, vol)
    with m.Worker(max_splats=16, max_cells=82, mesh_bytes=82 * 82 * 872) as w:
        blocks = w.marching_generate_host_field(vol, (0, 0, 0))
    assert len(blocks) > 1
    conv = [dict(vertices=b["vertices"], triangles=b["triangles"], num_internal=b["num_internal"],
                 keys=np.concatenate([np.zeros(b["num_internal"], np.uint64), b["external_keys"]])) for b in blocks]
    verts, tris = weld_by_key(conv)
    assert len(verts) == len(ref["vertices"]) and len(tris) == len(ref["triangles"])
    assert is_manifold(len(verts), tris) == ""
    # same vertex set
    a = np.unique(np.round(verts.astype(np.float64), 5), axis=0)
    b = np.unique(np.round(ref["vertices"].astype(np.float64), 5), axis=0)
    np.testing.assert_array_equal(a, b)


def test_empty_and_degenerate(orc):
    import mlsgpu_b200 as m
    with m.Worker(max_splats=16, max_cells=31) as w:
        assert w.marching_generate_host_field(np.ones((4, 5, 6), np.float32), (0, 0, 0)) == []
        assert w.marching_generate_host_field(np.full((4, 5, 6), np.nan, np.float32), (0, 0, 0)) == []
        assert w.marching_generate_host_field(-np.ones((1, 5, 6), np.float32), (0, 0, 0)) == []
        vol = sphere_field(12, 12, 12, 5.5, 5.5, 5.5, 3.2)
        vol[5, 2, 5] = np.nan
        blocks = w.marching_generate_host_field(vol, (0, 0, 0))
        assert_block_equal(blocks[0], oracle_mesh(orc, vol))
        with pytest.raises(m.MlsbError):
            w.marching_generate_host_field(np.ones((40, 5, 6), np.float32), (0, 0, 0))


def test_random_field(orc):
    import mlsgpu_b200 as m
    rng = np.random.default_rng(11)
    vol = rng.normal(size=(20, 37, 41)).astype(np.float32)
    vol[rng.uniform(size=vol.shape) < 0.02] = np.nan
    ref = oracle_mesh(orc, vol, (1, 2, 3))
    with m.Worker(max_splats=16, max_cells=63) as w:
        blocks = w.marching_generate_host_field(vol, (1, 2, 3))
    assert_block_equal(blocks[0], ref)


@pytest.mark.parametrize("width", [33, 65, 257, 34, 300, 512])
def test_row_widths_around_the_lattice_word_size(orc, width):
    """Cell rows of exactly 32, 64 and 256 cells (and 33): the vertex of the last corner column is carried
    into the extra lattice word of the row, and random NaN holes cut the surface everywhere.  300 and
    512 corners need the two-chunk variant of the cell kernels (bins up to 511 cells per side)."""
    import mlsgpu_b200 as m
    rng = np.random.RandomState(width)
    vol = rng.standard_normal((6, 11, width)).astype(np.float32)
    vol[rng.rand(*vol.shape) < 0.1] = np.nan
    vol[:, :, -1] = np.where(np.arange(11)[None, :] % 2 == 0, 1.0, -1.0)      # crossings in the last column
    ref = oracle_mesh(orc, vol, (3, 5, 7))
    with m.Worker(max_splats=16, max_cells=width - 1 if width > 32 else 32) as w:
        blocks = w.marching_generate_host_field(vol, (3, 5, 7))
    assert len(blocks) == 1 and len(ref["vertices"]) > ref["num_internal"] > 0
    assert_block_equal(blocks[0], ref)
    np.testing.assert_array_equal(blocks[0]["triangles"], ref["triangles"])
