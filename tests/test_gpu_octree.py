"""GPU octree build vs the oracle: start[] and commands[] bit-identical (integer path)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.oracle import Oracle
from fixtures import make_splats, walk_chain
from test_oracle_octree import BUILD_SPLATS, check_build, random_splats


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def gpu_build(splats, first, n, offset, subsampling, max_levels, size=None):
    import mlsgpu_b200 as m
    with m.Worker(max_levels=max_levels, subsampling=max(subsampling, 0), max_splats=max(len(splats), 16),
                  max_cells=31) as w:
        w.upload_splats(splats)
        w.octree_fill_commands(0xEF)
        if size is None:
            size = (1 << (max_levels + subsampling - 1),) * 3
        w.octree_build(first, n, size, offset, subsampling)
        start, commands = w.octree_get()
        out = w.download_splats(0, len(splats))
    return start, commands, out


def compare(orc, splats, first, n, offset, subsampling, max_levels):
    ref_splats = splats.copy()
    ref = orc.octree_build(ref_splats, first, n, offset, subsampling, max_levels, fill=np.int32(-269488145))  # 0xEFEFEFEF
    start, commands, out = gpu_build(splats, first, n, offset, subsampling, max_levels)
    np.testing.assert_array_equal(start, ref["start"])
    m = min(len(commands), len(ref["commands"]))
    np.testing.assert_array_equal(commands[:m], ref["commands"][:m])
    np.testing.assert_array_equal(out.view(np.uint32), ref_splats.view(np.uint32))
    return start, commands, ref


def test_build_reference_fixture(orc):
    # TestSplatTree::testBuild inputs (test/test_splat_tree.cpp:71-100): levels 9, subsampling 0
    splats = make_splats(BUILD_SPLATS)
    start, commands, ref = compare(orc, splats, 0, len(splats), (3, 0, 1), 0, 9)
    check_build(orc.make_code, splats, dict(start=start, commands=commands, used=ref["used"]), (16, 16, 12), (3, 0, 1))


def test_random_reference_fixture(orc):
    # TestSplatTree::testRandom inputs: 207 splats, subsampling 2, levels 8, negative offset
    splats, cells = random_splats()
    start, commands, _ = compare(orc, splats, 0, len(splats), (1, 2, -1), 2, 8)
    for z in range(0, cells[2] + 1, 4):
        for y in range(0, cells[1] + 1, 4):
            for x in range(0, cells[0] + 1, 4):
                found = walk_chain(start, commands, orc.make_code(x >> 2, y >> 2, z >> 2))
                assert all(0 <= c < len(splats) for c in found)


def test_empty_and_first_splat(orc):
    splats = make_splats([(100.0, 100.0, 100.0, 1.0)] * 3 + [(4.5, 4.5, 4.5, 1.0)])
    compare(orc, splats, 3, 1, (0, 0, 0), 3, 6)
    compare(orc, splats, 0, 0, (0, 0, 0), 3, 6)


@pytest.mark.parametrize("n,seed", [(5000, 1), (200000, 2)])
def test_default_levels_sphere(orc, n, seed):
    """CLI defaults (levels 6, subsampling 3) on a noisy sphere bin: multi-tile sort and scan."""
    from mlsgpu_b200 import host
    s = host.sphere_splats(n, seed=seed, radius=0.006)
    g, bins, lo, spacing = host.prepare(s, 0.01)
    assert len(bins) == 1
    b = bins[0]
    splats = np.ascontiguousarray(g[b["ids"]])
    compare(orc, splats, 0, len(splats), b["offset"], 3, 6)


def test_mixed_radii_levels(orc):
    """Radii spanning several octree levels, some splats outside the region."""
    rng = np.random.default_rng(7)
    n = 30000
    rows = np.zeros((n, 4), np.float32)
    rows[:, :3] = rng.uniform(-10, 270, (n, 3))
    rows[:, 3] = np.exp(rng.uniform(np.log(0.3), np.log(60.0), n))
    splats = make_splats([tuple(r) for r in rows])
    compare(orc, splats, 0, n, (5, 0, 9), 3, 6)


def test_large_bin_uses_large_sort_tiles(orc):
    """>= 3 M entry slots switch the radix sort to 16384-key tiles; results stay bit-identical."""
    from mlsgpu_b200 import host
    s = host.sphere_splats(500000, seed=8, radius=0.004)
    g, bins, lo, spacing = host.prepare(s, 0.01)
    assert len(bins) == 1
    b = bins[0]
    splats = np.ascontiguousarray(g[b["ids"]])
    compare(orc, splats, 0, len(splats), b["offset"], 3, 6)
