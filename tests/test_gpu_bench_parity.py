"""Parity ON THE WORKLOADS THE BENCH TIMES (BASELINE.json configs[0] and configs[1]): bins of the
C2 bumpy torus (per-splat radii, multi-level ancestor chains, batches of the MLS kernel that span
several command ranges) and the full-size C1 sphere, through the C ABI against the oracle.  Octree
arrays, field (NaN pattern and finite bits), pair-test / hit counts, welded vertices, external keys
and triangles must all be identical — the assertion style of test/test_mls.cpp:416-514 with the
tolerance at zero."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.oracle import Oracle
from fixtures import is_manifold
from gpu_common import (assert_block_equal, same_field, gpu_field, oracle_field, oracle_blocks, mesh_space, padded,
                        walk_chain)


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.fixture(scope="module")
def c2():
    """The bench workload itself: bench.make_workload(10 M, seed 2), 70 bins."""
    import bench
    return bench.make_workload(10_000_000, 2)


def pick_bins(wl):
    """largest, second largest, median, and the sparsest edge bin that still meshes."""
    order = sorted(range(len(wl["bins"])), key=lambda i: -wl["bins"][i]["n"])
    small = [i for i in order if 2000 <= wl["bins"][i]["n"] <= 30000]
    return {"largest": order[0], "second": order[1], "median": order[len(order) // 2],
            "sparse_edge": small[-1] if small else order[-1]}


def bin_splats(wl, b):
    return np.ascontiguousarray(wl["splats"][b["first"]:b["first"] + b["n"]])


@pytest.mark.parametrize("which", ["largest", "second", "median", "sparse_edge"])
def test_c2_bin_matches_oracle(orc, c2, which):
    import mlsgpu_b200 as m
    b = c2["bins"][pick_bins(c2)[which]]
    splats = bin_splats(c2, b)
    size, offset = b["size"], b["offset"]
    bias = c2["lo"].astype(np.float32) * c2["spacing"]
    mesh_bytes = 256 << 20                      # what bench.py gives every worker
    with m.Worker(max_splats=len(splats), max_cells=255, mesh_bytes=mesh_bytes) as w:
        w.set_scale_bias(float(c2["spacing"]), bias)
        got, start, commands, out, work = gpu_field(w, splats, size, offset)
        w.upload_splats(splats)
        blocks = w.process_bin(0, len(splats), size, offset)
    ref_img, tree, ref_splats, ref_work = oracle_field(orc, splats, size, offset)
    # octree: integer path, bit-exact; radii spanning several levels => multi-range chains
    np.testing.assert_array_equal(start, tree["start"])
    np.testing.assert_array_equal(commands[:tree["used"]], tree["commands"][:tree["used"]])
    np.testing.assert_array_equal(out.view(np.uint32), ref_splats.view(np.uint32))
    levels_used = {int(np.searchsorted(level_offsets(), k, side="right")) for k in np.flatnonzero(tree["start"] >= 0)[::97]}
    chain_ranges = max(count_ranges(tree["start"], tree["commands"], c) for c in np.flatnonzero(tree["start"][:32768] >= 0)[::53])
    if which != "sparse_edge":
        assert chain_ranges >= 2, "C2 bins must exercise chains of several command ranges"
    # field: the cull and the multi-range batches take exactly the oracle's decisions
    same_field(got, ref_img)
    assert work == ref_work
    assert np.isfinite(got).sum() > 1000
    # mesh
    vs, is_ = mesh_space(mesh_bytes)
    ref_blocks = oracle_blocks(orc, ref_img, size, offset, vs, is_)
    assert len(blocks) == len(ref_blocks) >= 1
    for g, r in zip(blocks, ref_blocks):
        assert_block_equal(g, r, float(c2["spacing"]), bias, exact=True)


def level_offsets(max_levels=6):
    off, pos = [], 0
    for i in range(max_levels):
        off.append(pos)
        pos += 1 << (3 * (max_levels - 1 - i))
    return np.array(off[1:])


def count_ranges(start, commands, code):
    """Number of command ranges on the fine -> root chain of a finest-level cell."""
    pos, n = int(start[code]), 0
    while pos >= 0:
        n += 1
        pos = int(commands[int(commands[pos])])
    return n


def test_c2_largest_bin_splits_at_64mb(orc, c2):
    """With the reference's default per-worker mesh memory the largest C2 bins overflow and are shipped in
    two blocks: each block equals the oracle's mesh of the same slice range, and the
    blocks weld by key into the mesh of the unsplit bin."""
    import mlsgpu_b200 as m
    b = c2["bins"][pick_bins(c2)["largest"]]
    splats = bin_splats(c2, b)
    size, offset = b["size"], b["offset"]
    mesh_bytes = 64 << 20
    with m.Worker(max_splats=len(splats), max_cells=255, mesh_bytes=mesh_bytes) as w:
        w.upload_splats(splats)
        blocks = w.process_bin(0, len(splats), size, offset)
    ref_img, _, _, _ = oracle_field(orc, splats, size, offset)
    vs, is_ = mesh_space(mesh_bytes)
    ref_blocks = oracle_blocks(orc, ref_img, size, offset, vs, is_)
    assert len(ref_blocks) >= 2 and len(blocks) == len(ref_blocks)
    for g, r in zip(blocks, ref_blocks):
        assert_block_equal(g, r, exact=True)
    whole = oracle_blocks(orc, ref_img, size, offset, 1 << 30, 1 << 30)[0]
    assert sum(len(g["triangles"]) for g in blocks) == len(whole["triangles"])


def test_c1_full_size_matches_oracle(orc):
    """BASELINE.json configs[0] at full size: 1 M-splat noisy sphere, fit-grid 0.01, ONE bin — a full
    compare with the oracle ('correctness config'), not properties."""
    import mlsgpu_b200 as m
    from mlsgpu_b200 import host
    s = host.sphere_splats(1_000_000, seed=1, radius=0.005, noise=0.002)
    g, bins, lo, spacing = host.prepare(s, 0.01)
    assert len(bins) == 1
    b = bins[0]
    splats = np.ascontiguousarray(g[b["ids"]])
    size, offset = b["size"], b["offset"]
    bias = lo.astype(np.float32) * spacing
    mesh_bytes = 256 << 20
    with m.Worker(max_splats=len(splats), max_cells=255, mesh_bytes=mesh_bytes) as w:
        w.set_scale_bias(float(spacing), bias)
        got, start, commands, out, work = gpu_field(w, splats, size, offset)
        w.upload_splats(splats)
        blocks = w.process_bin(0, len(splats), size, offset)
    ref_img, tree, ref_splats, ref_work = oracle_field(orc, splats, size, offset)
    np.testing.assert_array_equal(start, tree["start"])
    np.testing.assert_array_equal(commands[:tree["used"]], tree["commands"][:tree["used"]])
    np.testing.assert_array_equal(out.view(np.uint32), ref_splats.view(np.uint32))
    same_field(got, ref_img)
    assert work == ref_work
    vs, is_ = mesh_space(mesh_bytes)
    ref_blocks = oracle_blocks(orc, ref_img, size, offset, vs, is_)
    assert len(blocks) == len(ref_blocks) == 1
    assert_block_equal(blocks[0], ref_blocks[0], float(spacing), bias, exact=True)
    assert is_manifold(len(blocks[0]["vertices"]), blocks[0]["triangles"]) == ""


@pytest.mark.parametrize("shape", [1, 0], ids=["plane", "sphere"])
def test_mixed_radii_plane_fit_boundary_limit(orc, shape):
    """A bin whose radii span four octree levels (chains of up to four command ranges), plane fit and
    --fit-boundary-limit 0.5 on a rippled sheet with a hole and a ragged rim: the parts of the MLS
    kernel no reference test pins (test/test_mls.cpp:57,128) must still equal the oracle bit for bit."""
    import mlsgpu_b200 as m
    from mlsgpu_b200.capi import SPLAT_DTYPE
    rng = np.random.Generator(np.random.Philox(77))
    n = 120_000
    xy = rng.uniform(0.0, 1.0, (n, 2))
    keep = (np.hypot(xy[:, 0] - 0.55, xy[:, 1] - 0.4) > 0.12) & (xy[:, 0] + 0.15 * np.sin(25 * xy[:, 1]) < 0.93)
    xy = xy[keep]
    n = len(xy)
    z = 0.5 + 0.03 * np.sin(9 * xy[:, 0]) * np.cos(7 * xy[:, 1])
    nrm = np.stack([-0.27 * np.cos(9 * xy[:, 0]) * np.cos(7 * xy[:, 1]), 0.21 * np.sin(9 * xy[:, 0]) * np.sin(7 * xy[:, 1]),
                    np.ones(n)], axis=1)
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    s = np.zeros(n, SPLAT_DTYPE)
    s["position"] = np.stack([xy[:, 0], xy[:, 1], z + rng.normal(0, 0.0008, n)], axis=1).astype(np.float32)
    s["normal"] = nrm.astype(np.float32)
    # world radii x smooth 4 / grid 0.004 -> 2.2 ... 30 cells: level shifts 3, 4, 5 and 6
    s["radius"] = (0.0022 * np.exp(rng.uniform(0.0, np.log(13.5), n))).astype(np.float32)
    from mlsgpu_b200 import host
    g, bins, lo, spacing = host.prepare(s, 0.004)
    b = max(bins, key=lambda q: len(q["ids"]))
    splats = np.ascontiguousarray(g[b["ids"]])
    size, offset = b["size"], b["offset"]
    with m.Worker(max_splats=len(splats), max_cells=255, mesh_bytes=256 << 20, shape=shape, boundary_limit=0.5) as w:
        got, start, commands, out, work = gpu_field(w, splats, size, offset)
        w.upload_splats(splats)
        blocks = w.process_bin(0, len(splats), size, offset)
    ref_img, tree, ref_splats, ref_work = oracle_field(orc, splats, size, offset, limit=0.5, shape=shape)
    np.testing.assert_array_equal(start, tree["start"])
    np.testing.assert_array_equal(commands[:tree["used"]], tree["commands"][:tree["used"]])
    used_levels = np.unique(np.searchsorted(level_offsets(), np.flatnonzero(
        np.bincount(tree["keys"][tree["keys"] != 0xFFFFFFFF], minlength=37449)), side="right"))
    assert len(used_levels) >= 3, used_levels
    assert max(count_ranges(tree["start"], tree["commands"], c) for c in np.flatnonzero(tree["start"][:32768] >= 0)[::31]) >= 3
    same_field(got, ref_img)
    assert work == ref_work
    # the boundary test rejects corners the unlimited fit accepts
    open_img, _, _, _ = oracle_field(orc, splats, size, offset, limit=100.0, shape=shape)
    assert np.isfinite(open_img).sum() > np.isfinite(ref_img).sum() > 1000
    ref_blocks = oracle_blocks(orc, ref_img, size, offset, *mesh_space(256 << 20))
    assert len(blocks) == len(ref_blocks) == 1
    assert_block_equal(blocks[0], ref_blocks[0], exact=True)
