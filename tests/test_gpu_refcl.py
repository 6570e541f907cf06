"""The oracle (and, through it, the CUDA path) pinned against the REFERENCE'S OWN OpenCL kernels.

oracle/_ref/librefcl.so embeds kernels/octree.cl, mls.cl, marching.cl and scale_bias.cl of the
reference (built where /root/reference exists, see oracle/Makefile) behind our C host on the OpenCL C
API; on the GPU box NVIDIA's OpenCL runs them on the same B200.  The reference's host classes cannot
be built (Boost, clogs), so clogs' sort/scan are replaced by host code of the same semantics.

What this pins that no reference test does (test/test_mls.cpp:57,128): the plane fit, the
boundary test, vertex positions out of marching, and the rounding of the moment accumulation —
mls.cl leaves contraction to the OpenCL compiler, the oracle fuses.  Tolerances are the
north_star's: field 1e-5 (abs or rel, the reference's own MLSGPU_ASSERT_DOUBLES_EQUAL), positions
1e-4 grid units, integer results exact.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.oracle import Oracle
from oracle import refcl as refcl_mod
from gpu_common import canonical_triangles, oracle_field, padded


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.fixture(scope="module")
def ref():
    r = refcl_mod.get()
    if r is None:
        pytest.skip("reference OpenCL kernels unavailable: " + str(refcl_mod.why()))
    return r


def sheet_bin(seed=77, n=60_000, grid=0.006):
    """A rippled sheet with a hole, radii over several octree levels (chains of several ranges)."""
    from mlsgpu_b200.capi import SPLAT_DTYPE
    from mlsgpu_b200 import host
    rng = np.random.Generator(np.random.Philox(seed))
    xy = rng.uniform(0.0, 1.0, (n, 2))
    xy = xy[np.hypot(xy[:, 0] - 0.55, xy[:, 1] - 0.4) > 0.12]
    n = len(xy)
    z = 0.5 + 0.03 * np.sin(9 * xy[:, 0]) * np.cos(7 * xy[:, 1])
    nrm = np.stack([-0.27 * np.cos(9 * xy[:, 0]) * np.cos(7 * xy[:, 1]), 0.21 * np.sin(9 * xy[:, 0]) * np.sin(7 * xy[:, 1]),
                    np.ones(n)], axis=1)
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    s = np.zeros(n, SPLAT_DTYPE)
    s["position"] = np.stack([xy[:, 0], xy[:, 1], z + rng.normal(0, 0.0008, n)], axis=1).astype(np.float32)
    s["normal"] = nrm.astype(np.float32)
    s["radius"] = (0.003 * np.exp(rng.uniform(0.0, np.log(8.0), n))).astype(np.float32)
    g, bins, lo, spacing = host.prepare(s, grid)
    b = max(bins, key=lambda q: len(q["ids"]))
    return np.ascontiguousarray(g[b["ids"]]), b["size"], b["offset"]


def test_device_is_the_gpu(ref):
    assert "NVIDIA" in ref.device or "B200" in ref.device, ref.device


def test_octree_build_reference_kernels_vs_oracle(orc, ref):
    """writeEntries / countCommands / writeSplatIds / writeStart[Top] / fill of kernels/octree.cl vs
    the oracle: entry keys, start[] and commands[] are integer results and must be identical."""
    splats, size, offset = sheet_bin()
    a, b = splats.copy(), splats.copy()
    got = ref.octree_build(a, 0, len(a), offset, 3, 6)
    want = orc.octree_build(b, 0, len(b), offset, 3, 6)
    np.testing.assert_array_equal(got["keys"], want["keys"])
    np.testing.assert_array_equal(got["values"], want["values"])
    np.testing.assert_array_equal(got["start"], want["start"])
    np.testing.assert_array_equal(got["commands"][:want["used"]], want["commands"][:want["used"]])
    # radius <- 1/r^2: OpenCL's default division is within 2.5 ulp, the oracle's is IEEE
    np.testing.assert_allclose(a["radius"], b["radius"], rtol=4e-7)
    np.testing.assert_array_equal(a["position"], b["position"])


def compare_fields(got, want, tol=1e-5):
    """NaN pattern and values: returns (corners whose validity differs, finite corners in both, max error)."""
    ng, nw = np.isnan(got), np.isnan(want)
    both = ~ng & ~nw
    err = np.abs(got[both].astype(np.float64) - want[both].astype(np.float64))
    err = np.minimum(err, err / np.maximum(np.abs(want[both].astype(np.float64)), 1e-30))   # abs or rel
    return int((ng != nw).sum()), int(both.sum()), float(err.max()) if err.size else 0.0, \
        int((got[both].view(np.uint32) == want[both].view(np.uint32)).sum())


@pytest.mark.parametrize("shape", [0, 1], ids=["sphere", "plane"])
@pytest.mark.parametrize("limit", [1.0, 0.5])
def test_process_corners_reference_kernel_vs_oracle(orc, ref, shape, limit):
    """processCorners of kernels/mls.cl (sphere AND plane fit, boundary limit 1.0 AND 0.5) on the
    oracle's octree vs the oracle's field: every corner within 1e-5; validity may differ only on a
    handful of corners whose predicates sit within rounding of their thresholds (enumerated)."""
    splats, size, offset = sheet_bin()
    want, tree, mutated, _ = oracle_field(orc, splats, size, offset, limit=limit, shape=shape)
    pitch, zs, depth = padded(size)
    got = np.zeros((zs * depth, pitch), np.float32)
    secs, launches = ref.mls_process_corners(got, mutated, tree["commands"][:tree["used"] + 1], tree["start"], 9, offset,
                                             int(size[0]), int(size[1]), zs, 0, 0, int(size[2]) - 1,
                                             orc.boundary_factor(limit), shape)
    flips, both, err, identical = compare_fields(got, want)
    print("shape %d limit %.1f: %d finite corners, %d bit-identical, max err %.3g, %d validity flips, %.3f ms in %d launches"
          % (shape, limit, both, identical, err, flips, 1e3 * secs, launches))
    assert both > 20_000
    assert err < 1e-5
    assert flips <= max(3, both // 5000), "validity-ambiguous corners: %d of %d" % (flips, both)
    if limit == 0.5:
        # the boundary test is live: it rejects corners the open fit accepts, in both implementations alike
        open_, _, _, _ = oracle_field(orc, splats, size, offset, limit=100.0, shape=shape)
        assert np.isfinite(open_).sum() > np.isfinite(want).sum() + 1000


def test_process_corners_swathes_agree(orc, ref):
    """The same field whether the reference kernel is driven in one image or in 8-slice swathes."""
    splats, size, offset = sheet_bin(n=20_000, grid=0.012)
    want, tree, mutated, _ = oracle_field(orc, splats, size, offset)
    pitch, zs, depth = padded(size)
    a = np.zeros((zs * depth, pitch), np.float32)
    b = np.zeros_like(a)
    args = (mutated, tree["commands"][:tree["used"] + 1], tree["start"], 9, offset, int(size[0]), int(size[1]), zs, 0, 0,
            int(size[2]) - 1, orc.boundary_factor(1.0), 0)
    ref.mls_process_corners(a, *args)
    _, launches = ref.mls_process_corners(b, *args, max_swathe=8)
    assert launches == depth // 8
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))


def test_marching_reference_kernels_vs_oracle(orc, ref):
    """genOccupied / generateElements / countUniqueVertices / compactVertices / reindex /
    scaleBiasVertices on the oracle's field vs the oracle's mesh: counts, keys, internal/external
    split and connectivity identical; positions within 1e-4 grid units (scaled)."""
    splats, size, offset = sheet_bin()
    img, _, _, _ = oracle_field(orc, splats, size, offset)
    zs = padded(size)[1]
    tables = orc.marching_tables()
    ko = [int(v) for v in offset]
    scale, bias = 0.006, (0.25, -1.0, 3.0)
    got = ref.marching_generate(img, int(size[0]), int(size[1]), zs, 0, 0, int(size[2]) - 1, ko, tables, scale, bias)
    want = orc.marching_generate(img, int(size[0]), int(size[1]), zs, 0, 0, int(size[2]) - 1, ko)
    assert len(want["vertices"]) > 20_000
    assert got["stats"] == want["stats"]                       # occupied cells, unwelded vertices, indices
    assert len(got["vertices"]) == len(want["vertices"]) and got["num_internal"] == want["num_internal"]
    ni = want["num_internal"]
    np.testing.assert_array_equal(got["keys"][ni:], want["keys"][ni:])
    wv = want["vertices"].copy()
    orc.scale_bias(wv, scale, bias)
    err = np.abs(got["vertices"].astype(np.float64) - wv.astype(np.float64)).max()
    same = (got["vertices"].view(np.uint32) == wv.view(np.uint32)).all(axis=1).mean()
    print("marching: %d vertices, %.4f bit-identical, max position error %.3g world units (1e-4 grid units = %.3g)"
          % (len(wv), same, err, 1e-4 * scale))
    assert err <= 1e-4 * scale + 1e-7
    # the reference compacts cells with an atomic counter: triangle order is arbitrary, vertex order is by key
    np.testing.assert_array_equal(canonical_triangles(got["triangles"]), canonical_triangles(want["triangles"]))


def test_solve_quadratic_reference_kernel(orc, ref):
    """A few rows of the reference's own table (test/test_mls.cpp:287-347) through its UNIT_TESTS shim."""
    for a, b, c in [(1.0, 3.0, -4.0), (1e-20, 2.0, -3.0), (0.0, 5.0, 1e20), (1.0, 0.0, -1.0), (-1.0, 2.0, 3.0)]:
        g, w = ref.solve_quadratic(a, b, c), orc.solve_quadratic(a, b, c)
        if np.isnan(w):
            assert np.isnan(g)
        else:
            assert abs(g - w) <= 4 * np.finfo(np.float32).eps * max(1.0, abs(w)), (a, b, c, g, w)
