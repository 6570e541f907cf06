"""Committed golden vectors (tests/golden/make_golden.py): the oracle must reproduce them bit
for bit on any host, and the CUDA path must match them on the GPU."""
import os

import numpy as np
import pytest

from oracle.oracle import Oracle
from test_oracle_mls import process_corners_case
from test_oracle_marching import sphere_field, to_image
from test_oracle_octree import random_splats

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def test_oracle_reproduces_golden_field():
    orc = Oracle()
    c = process_corners_case(orc, 35.0)
    orc.mls_process_corners(c["image"], c["splats"], c["commands"], c["start"], 3 * c["sub"], c["offset"],
                            c["size"][0], c["size"][1], c["z_stride"], c["z_bias"], c["z_first"], c["z_last"],
                            orc.boundary_factor(1.0), 0)
    rows = slice(c["z_first"] * c["z_stride"] + c["z_bias"], (c["z_last"] + 1) * c["z_stride"] + c["z_bias"])
    np.testing.assert_array_equal(c["image"][rows].view(np.uint32), load("mls_process_corners.npz")["field"])


def test_oracle_reproduces_golden_octree():
    g = load("octree_random.npz")
    splats, _ = random_splats()
    t = Oracle().octree_build(splats, 0, len(splats), (1, 2, -1), subsampling=2, max_levels=8)
    np.testing.assert_array_equal(t["start"], g["start"])
    np.testing.assert_array_equal(t["commands"][:t["used"]], g["commands"])
    np.testing.assert_array_equal(splats["radius"].view(np.uint32), g["inv_r2"])


def test_oracle_reproduces_golden_mesh():
    g = load("marching_sphere.npz")
    vol = sphere_field(40, 36, 33, 19.5, 17.0, 16.0, 21.0)
    img, zs = to_image(vol)
    m = Oracle().marching_generate(img, 40, 36, zs, 0, 0, 32, (7, 100, 3))
    np.testing.assert_array_equal(m["vertices"].view(np.uint32), g["vertices"])
    np.testing.assert_array_equal(m["keys"][m["num_internal"]:], g["keys"])
    np.testing.assert_array_equal(m["triangles"], g["triangles"])
    assert m["num_internal"] == int(g["num_internal"][0])


@pytest.mark.gpu
def test_gpu_matches_golden_mesh():
    import mlsgpu_b200 as m
    g = load("marching_sphere.npz")
    vol = sphere_field(40, 36, 33, 19.5, 17.0, 16.0, 21.0)
    with m.Worker(max_splats=16, max_cells=63) as w:
        blocks = w.marching_generate_host_field(vol, (7, 100, 3))
    b = blocks[0]
    np.testing.assert_array_equal(b["vertices"].view(np.uint32), g["vertices"])
    np.testing.assert_array_equal(b["external_keys"], g["keys"])
    np.testing.assert_array_equal(b["triangles"], g["triangles"])
    assert b["num_internal"] == int(g["num_internal"][0])


@pytest.mark.gpu
def test_gpu_matches_golden_octree():
    import mlsgpu_b200 as m
    g = load("octree_random.npz")
    splats, _ = random_splats()
    with m.Worker(max_levels=8, subsampling=2, max_splats=len(splats), max_cells=31) as w:
        w.upload_splats(splats)
        w.octree_build(0, len(splats), (512, 512, 512), (1, 2, -1), 2)
        start, commands = w.octree_get(len(g["commands"]))
        out = w.download_splats(0, len(splats))
    np.testing.assert_array_equal(start, g["start"])
    np.testing.assert_array_equal(commands, g["commands"])
    np.testing.assert_array_equal(out["radius"].view(np.uint32), g["inv_r2"])
