"""One bin end to end through mlsb_worker_process_bin (DeviceWorkerGroupBase::Worker::operator())
against the oracle chained the same way; multi-bin meshes weld into a manifold."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.oracle import Oracle
from fixtures import is_manifold
from gpu_common import assert_block_equal
from test_oracle_marching import weld_by_key


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def oracle_bin(orc, splats, offset, size, spacing, lo, limit=1.0, shape=0):
    s = splats.copy()
    tree = orc.octree_build(s, 0, len(s), offset, 3, 6)
    pitch = -(-int(size[0]) // 8) * 8
    zs = -(-int(size[1]) // 8) * 8
    depth = -(-int(size[2]) // 8) * 8
    img = np.zeros((zs * depth, pitch), np.float32)
    orc.mls_process_corners(img, s, tree["commands"], tree["start"], 9, offset, int(size[0]), int(size[1]), zs, 0, 0,
                            int(size[2]) - 1, orc.boundary_factor(limit), shape)
    return orc.marching_generate(img, int(size[0]), int(size[1]), zs, 0, 0, int(size[2]) - 1, [int(v) for v in offset])


def test_single_bin_sphere(orc):
    import mlsgpu_b200 as m
    from mlsgpu_b200 import host
    s = host.sphere_splats(80000, seed=1, radius=0.01, noise=0.003)
    g, bins, lo, spacing = host.prepare(s, 0.02)
    assert len(bins) == 1
    b = bins[0]
    splats = np.ascontiguousarray(g[b["ids"]])
    bias = lo.astype(np.float32) * spacing
    ref = oracle_bin(orc, splats, b["offset"], b["size"], spacing, lo)
    with m.Worker(max_splats=len(splats), max_cells=int(b["size"].max()) - 1) as w:
        w.set_scale_bias(float(spacing), bias)
        w.upload_splats(splats)
        blocks = w.process_bin(0, len(splats), b["size"], b["offset"])
        assert w.launch_count > 0
    assert len(blocks) == 1
    assert_block_equal(blocks[0], ref, float(spacing), bias)
    v = blocks[0]["vertices"]
    assert is_manifold(len(v), blocks[0]["triangles"]) == ""
    # closed surface close to the unit sphere
    assert blocks[0]["num_internal"] == len(v)
    assert np.abs(np.linalg.norm(v, axis=1) - 1.0).max() < 0.02


def test_multi_bin_welds_to_manifold(orc):
    import mlsgpu_b200 as m
    from mlsgpu_b200 import host
    s = host.sphere_splats(60000, seed=4, radius=0.012, noise=0.002)
    g, bins, lo, spacing = host.prepare(s, 0.025, max_cells=47)
    assert len(bins) >= 8
    out = []
    with m.Worker(max_splats=len(g), max_cells=47) as w:
        for b in bins:
            splats = np.ascontiguousarray(g[b["ids"]])
            w.upload_splats(splats)
            blocks = w.process_bin(0, len(splats), b["size"], b["offset"])
            ref = oracle_bin(orc, splats, b["offset"], b["size"], spacing, lo)
            if len(ref["vertices"]) == 0:
                assert blocks == []
                continue
            assert_block_equal(blocks[0], ref)
            out.append(dict(vertices=ref["vertices"], triangles=ref["triangles"], num_internal=ref["num_internal"],
                            keys=ref["keys"]))
    verts, tris = weld_by_key(out)
    assert is_manifold(len(verts), tris) == ""


def test_worker_group_matches_single_worker(orc):
    """CopyGroup + DeviceWorkerGroup equivalent: bins submitted to 3 native worker threads give, per
    bin, exactly the blocks a single worker produces; the welded result is a manifold."""
    import mlsgpu_b200 as m
    from mlsgpu_b200 import host
    s = host.sphere_splats(60000, seed=4, radius=0.012, noise=0.002)
    g, bins, lo, spacing = host.prepare(s, 0.025, max_cells=47)
    got = {}

    def sink(tag, block):
        got.setdefault(tag, []).append(block)

    max_n = max(len(b["ids"]) for b in bins)
    with m.WorkerGroup(num_workers=3, sink=sink, max_splats=max_n, max_cells=47) as grp:
        for i, b in enumerate(bins):
            grp.submit(np.ascontiguousarray(g[b["ids"]]), b["size"], b["offset"], tag=i)   # pageable memory: staged
        grp.wait()
        nb, ns, mesh_bytes = grp.counters()
        assert nb == len(bins) and ns == sum(len(b["ids"]) for b in bins) and mesh_bytes > 0
        assert grp.released() == len(bins)      # every bin's splat memory has been handed back
        # an oversized bin is reported by wait(), not swallowed
        with pytest.raises(m.MlsbError):
            grp.submit(np.ascontiguousarray(g[bins[0]["ids"]]), (400, 8, 8), bins[0]["offset"], tag=99)
            grp.wait()
    out = []
    with m.Worker(max_splats=max_n, max_cells=47) as w:
        for i, b in enumerate(bins):
            w.upload_splats(np.ascontiguousarray(g[b["ids"]]))
            blocks = w.process_bin(0, len(b["ids"]), b["size"], b["offset"])
            assert len(blocks) == len(got.get(i, []))
            for a, c in zip(blocks, got.get(i, [])):
                np.testing.assert_array_equal(a["vertices"], c["vertices"])
                np.testing.assert_array_equal(a["triangles"], c["triangles"])
                np.testing.assert_array_equal(a["external_keys"], c["external_keys"])
                ni = a["num_internal"]
                out.append(dict(vertices=a["vertices"], triangles=a["triangles"], num_internal=ni,
                                keys=np.concatenate([np.zeros(ni, np.uint64), a["external_keys"]])))
    verts, tris = weld_by_key(out)
    assert is_manifold(len(verts), tris) == ""


def test_c1_full_size_properties():
    """BASELINE.json configs[0] at full size (1 M-splat noisy sphere, fit-grid 0.01, single bin),
    checked through size-independent properties: one shipped block, oriented manifold, F ~ 2V,
    vertices on the sphere, keys absent (no external vertices), and the
    run is reproducible bit for bit (also through the worker group)."""
    import mlsgpu_b200 as m
    from mlsgpu_b200 import host
    s = host.sphere_splats(1_000_000, seed=1, radius=0.005, noise=0.002)
    g, bins, lo, spacing = host.prepare(s, 0.01)
    assert len(bins) == 1
    b = bins[0]
    assert b["size"].max() <= 256 and len(b["ids"]) == 1_000_000
    splats = np.ascontiguousarray(g[b["ids"]])
    bias = lo.astype(np.float32) * spacing
    runs = []
    with m.Worker(max_splats=len(splats), max_cells=255, mesh_bytes=256 << 20) as w:
        w.set_scale_bias(float(spacing), bias)
        for _ in range(2):
            w.upload_splats(splats)
            runs.append(w.process_bin(0, len(splats), b["size"], b["offset"]))
    got = []
    with m.WorkerGroup(num_workers=2, sink=lambda tag, blk: got.append(blk), max_splats=len(splats), max_cells=255,
                       mesh_bytes=256 << 20) as grp:
        for w in grp.workers:
            w.set_scale_bias(float(spacing), bias)
        grp.submit(splats, b["size"], b["offset"])
        grp.wait()
    assert len(runs[0]) == 1 and len(got) == 1
    blk = runs[0][0]
    for other in (runs[1][0], got[0]):
        np.testing.assert_array_equal(blk["vertices"].view(np.uint32), other["vertices"].view(np.uint32))
        np.testing.assert_array_equal(blk["triangles"], other["triangles"])
    v, t = blk["vertices"], blk["triangles"]
    assert blk["num_internal"] == len(v) and len(blk["external_keys"]) == 0
    assert np.abs(np.linalg.norm(v.astype(np.float64), axis=1) - 1.0).max() < 0.01
    assert len(v) > 100_000 and abs(len(t) - 2 * len(v)) < 0.01 * len(v)   # F ~ 2V: closed up to a few small holes
    assert is_manifold(len(v), t) == ""


def test_upload_from_device_memory():
    """mlsb_splats_upload_async with a source in DEVICE memory (copied by a kernel, not by the copy engine): the
    worker's buffer holds the same bytes as after an upload from the host, at an offset on both sides, for a
    16-byte-aligned source and for one that is not (the cudaMemcpyAsync fall-back)."""
    import mlsgpu_b200 as m
    from mlsgpu_b200 import capi, host
    s = host.sphere_splats(5003, seed=3, radius=0.02, noise=0.004)
    dev = capi.DeviceBuffer.from_array(s)
    raw = np.concatenate([np.zeros(4, np.uint8), s.view(np.uint8)])        # the splats start 4 bytes into this one
    odd = capi.DeviceBuffer.from_array(raw)
    with m.Worker(max_splats=6000, max_cells=31) as w:
        launches = w.launch_count
        w.upload_splats_from_device(dev, 7, 4990, dst_first=11)
        assert w.launch_count == launches + 1
        np.testing.assert_array_equal(w.download_splats(11, 4990).view(np.uint8), s[7:4997].view(np.uint8))
        from mlsgpu_b200.capi import lib, _check
        _check(lib().mlsb_splats_upload_async(w.h, odd.ptr.value + 4, 0, len(s)))
        assert w.launch_count == launches + 1                               # not 16-byte aligned: the copy engine took it
        np.testing.assert_array_equal(w.download_splats(0, len(s)).view(np.uint8), s.view(np.uint8))
    dev.free()
    odd.free()
