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
