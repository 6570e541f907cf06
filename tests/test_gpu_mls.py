"""GPU processCorners vs the oracle: field bit-identical, plus the reference's own
known-answer test (test/test_mls.cpp:416-514) run directly on the CUDA kernel."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.oracle import Oracle
from test_oracle_mls import process_corners_case, check_process_corners


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def same_field(a, b):
    """NaN positions equal and finite values bit-identical."""
    na, nb = np.isnan(a), np.isnan(b)
    np.testing.assert_array_equal(na, nb)
    np.testing.assert_array_equal(a[~na].view(np.uint32), b[~nb].view(np.uint32))


@pytest.mark.parametrize("radius", [65.0, 35.0])
def test_process_corners_reference_fixture(orc, radius):
    import mlsgpu_b200 as m
    from mlsgpu_b200.capi import DeviceBuffer
    c = process_corners_case(orc, radius)
    ref = c["image"].copy()
    orc.mls_process_corners(ref, c["splats"], c["commands"], c["start"], 3 * c["sub"], c["offset"], c["size"][0],
                            c["size"][1], c["z_stride"], c["z_bias"], c["z_first"], c["z_last"],
                            orc.boundary_factor(1.0), 0)
    with m.Worker(max_splats=64, max_cells=31, subsampling=c["sub"]) as w:
        d_splats = DeviceBuffer.from_array(c["splats"])
        d_cmd = DeviceBuffer.from_array(c["commands"])
        d_start = DeviceBuffer.from_array(c["start"])
        d_img = DeviceBuffer.from_array(c["image"])
        w.mls_set_buffers(c["offset"], d_splats, d_cmd, d_start, c["sub"])
        w.mls_corners(d_img, c["image"].shape[1], c["size"][0], c["size"][1], c["z_stride"], c["z_bias"],
                      c["z_first"], c["z_last"])
        w.synchronize()
        got = d_img.to_array(np.float32, c["image"].shape)
    check_process_corners(c, got)          # the reference's own assertions (analytic sphere, 1e-5)
    same_field(got, ref)                   # and bit-identical to the oracle
    # rows below the swathe are untouched (post-condition, src/marching.h:243-246)
    first_row = c["z_first"] * c["z_stride"] + c["z_bias"]
    assert (got[:first_row] == np.float32(12345.0)).all()


@pytest.mark.parametrize("shape", [0, 1], ids=["sphere", "plane"])
@pytest.mark.parametrize("limit", [1.0, 0.5])
def test_field_matches_oracle_on_sphere_bin(orc, shape, limit):
    """Octree from the GPU build, field from the GPU MLS kernel vs the oracle on the same octree."""
    import mlsgpu_b200 as m
    from mlsgpu_b200 import host
    from mlsgpu_b200.capi import DeviceBuffer
    s = host.sphere_splats(60000, seed=5, radius=0.012, noise=0.004)
    # cut a hole so that the boundary test has something to reject
    s = s[~((s["position"][:, 2] > 0.8) & (s["position"][:, 0] > 0))]
    g, bins, lo, spacing = host.prepare(s, 0.025)
    assert len(bins) == 1
    b = bins[0]
    splats = np.ascontiguousarray(g[b["ids"]])
    size = b["size"]
    with m.Worker(max_splats=len(splats), max_cells=int(size.max()) - 1, shape=shape, boundary_limit=limit) as w:
        w.set_profiling(True)
        w.set_work_counting(True)
        w.upload_splats(splats)
        w.octree_build(0, len(splats), [-(-int(v) // 8) * 8 for v in size], b["offset"])
        w.mls_set(b["offset"])
        pitch = -(-int(size[0]) // 8) * 8
        zs = -(-int(size[1]) // 8) * 8
        depth = -(-int(size[2]) // 8) * 8
        d_img = DeviceBuffer(4 * pitch * zs * depth)
        d_img.fill(0)
        w.mls_corners(d_img, pitch, int(size[0]), int(size[1]), zs, 0, 0, int(size[2]) - 1)
        w.synchronize()
        got = d_img.to_array(np.float32, (zs * depth, pitch))
        start, commands = w.octree_get()
        dsplats = w.download_splats(0, len(splats))
        work = w.mls_work()
        secs, launches = w.statistic("kernel.mls.processCorners.time")
    ref = np.zeros_like(got)
    t, h = orc.mls_process_corners(ref, dsplats, commands, start, 9, b["offset"], int(size[0]), int(size[1]), zs, 0, 0,
                                   int(size[2]) - 1, orc.boundary_factor(limit), shape)
    same_field(got, ref)
    assert np.isfinite(got).sum() > 1000
    assert work == (t, h) and launches == 1 and secs > 0


def test_hit_threshold_boundary_cases(orc):
    """Splats whose 0.99 cut-off falls within a few ulp of a corner's |p|^2, plus degenerate
    radii (1/r^2 = 0, inf, NaN, denormal, huge): the kernel's threshold compare must take exactly
    the decisions of `pp * invr2 < 0.99f` — same hit count, bit-identical field."""
    import mlsgpu_b200 as m
    from mlsgpu_b200.capi import DeviceBuffer, SPLAT_DTYPE
    rng = np.random.default_rng(3)
    n = 600
    splats = np.zeros(n, SPLAT_DTYPE)
    splats["position"] = rng.uniform(0.0, 16.0, (n, 3)).astype(np.float32)
    nrm = rng.normal(size=(n, 3))
    splats["normal"] = (nrm / np.linalg.norm(nrm, axis=1)[:, None]).astype(np.float32)
    splats["quality"] = rng.uniform(0.5, 2.0, n).astype(np.float32)
    inv = np.zeros(n, np.float32)
    for i in range(n):
        c = rng.integers(0, 16, 3).astype(np.float32)           # a corner of the region
        p = splats["position"][i] - c
        pp = np.float32(np.float32(p[0] * p[0]) + np.float32(p[1] * p[1]) + np.float32(p[2] * p[2]))
        w = np.float32(0.99) / max(pp, np.float32(1e-3))
        inv[i] = np.nextafter(w, np.float32(np.inf if rng.integers(2) else 0), dtype=np.float32) if rng.integers(3) else w
        for _ in range(int(rng.integers(0, 3))):
            inv[i] = np.nextafter(inv[i], np.float32(np.inf if rng.integers(2) else 0), dtype=np.float32)
    inv[:8] = np.array([0.0, np.inf, np.nan, 1e-45, 1e-38, 3e38, 1e30, 1e-30], np.float32)
    splats["radius"] = inv     # the kernel expects 1/r^2 here
    commands = np.array([n + 1] + list(range(n)) + [-1], np.int32)
    start = np.zeros(8, np.int32)
    size, sub = (16, 16, 16), 3
    img = np.zeros((16 * 16, 16), np.float32)
    ref = img.copy()
    t, h = orc.mls_process_corners(ref, splats, commands, start, 3 * sub, (0, 0, 0), 16, 16, 16, 0, 0, 15,
                                   orc.boundary_factor(1.0), 0)
    with m.Worker(max_splats=64, max_cells=31) as w:
        w.set_work_counting(True)
        bufs = [DeviceBuffer.from_array(a) for a in (splats, commands, start, img)]
        w.mls_set_buffers((0, 0, 0), bufs[0], bufs[1], bufs[2], sub)
        w.mls_corners(bufs[3], 16, 16, 16, 16, 0, 0, 15)
        w.synchronize()
        got = bufs[3].to_array(np.float32, img.shape)
        assert w.mls_work() == (t, h)
    same_field(got, ref)
    assert h > 1000
