"""Pins the octree oracle against the reference's known-answer tests
(test/test_splat_tree_cl.cpp:170-204, test/test_splat_tree.cpp:46-244)."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from fixtures import MT19937, make_splats, walk_chain


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def test_level_shift(orc):
    # test/test_splat_tree_cl.cpp:170-182
    cases = [((0, 0, 0), (0, 0, 0), 0), ((1, 1, 1), (0, 0, 0), 0), ((0, 1, 2), (1, 2, 3), 0),
             ((0, 1, 2), (2, 2, 3), 1), ((0, 1, 2), (1, 3, 3), 1), ((0, 1, 2), (1, 2, 4), 1),
             ((31, 0, 0), (35, 0, 0), 2), ((31, 0, 0), (36, 0, 0), 3), ((27, 0, 0), (32, 0, 0), 3),
             ((48, 0, 0), (79, 0, 0), 5)]
    for lo, hi, want in cases:
        assert orc.level_shift(lo, hi) == want


def test_point_box_dist2(orc):
    # test/test_splat_tree_cl.cpp:184-196
    assert orc.point_box_dist2((0.5, 0.5, 0.5), (0, 0, 0), (1, 1, 1)) == pytest.approx(0.0, abs=1e-4)
    assert orc.point_box_dist2((0.25, 0.5, 3.0), (-1.5, 0.0, 0.5), (1.5, 0.75, 1.0)) == pytest.approx(4.0, abs=1e-4)
    assert orc.point_box_dist2((9.0, 11.0, -10.0), (-1.0, 0.0, -7.0), (8.0, 9.0, 8.0)) == pytest.approx(14.0, abs=1e-4)


def test_make_code(orc):
    # test/test_splat_tree_cl.cpp:198-204 and test/test_splat_tree.cpp:46-54
    assert orc.make_code(0, 0, 0) == 0
    assert orc.make_code(1, 1, 1) == 7
    assert orc.make_code(2, 5, 3) == 174
    assert orc.make_code(7, 7, 7) == 511
    assert orc.make_code(123, 456, 789) == 642569997


BUILD_SPLATS = [(10.5, 7.5, 8.5, 1.0), (11.5, 9.5, 5.5, 2.0), (10.0, 10.0, 6.0, 4.5), (3.0, 0.5, 1.0, 0.75),
                (3.0, 0.0, 1.0, 0.5),
                # spheres lying entirely outside the octree
                (19.0, 8.0, 5.0, 0.6), (0.0, 1.0, 1.0, 2.5), (5.0, 1000.0, 5.0, 900.0)]


def check_build(orc_make_code, splats_in, result, size, offset):
    """The walk checks of TestSplatTree::testBuild, test/test_splat_tree.cpp:71-170."""
    start, commands = result["start"], result["commands"]
    assert 16 * 16 * 16 <= len(start)
    max_pos = -1
    for z in range(size[2]):
        for y in range(size[1]):
            for x in range(size[0]):
                idx = orc_make_code(x, y, z)
                pos = int(start[idx])
                max_pos = max(max_pos, pos)
                found = walk_chain(start, commands, idx, limit=1000)
                assert all(0 <= c < len(splats_in) for c in found)
                assert len(set(found)) == len(found), "splat listed twice"
                corner = np.array([x + offset[0], y + offset[1], z + offset[2]], np.float32)
                for i, s in enumerate(splats_in):
                    p = s["position"]
                    n = np.maximum(np.minimum(p, corner + np.float32(1.0)), corner) - p
                    dist2 = np.float32(0)
                    for j in range(3):
                        dist2 += n[j] * n[j]
                    if dist2 <= s["radius"] * s["radius"]:
                        assert i in found, (x, y, z, i)
    # no splat in more than 8 ranges
    used = result["used"]
    repeats = {}
    i = 0
    while i < used:
        end = int(commands[i])
        i += 1
        while i < end:
            c = int(commands[i])
            repeats[c] = repeats.get(c, 0) + 1
            assert repeats[c] <= 8
            i += 1
        i += 1


def test_build(orc):
    size = (16, 16, 12)
    offset = (3, 0, 1)
    splats = make_splats(BUILD_SPLATS)
    orig = splats.copy()
    res = orc.octree_build(splats, 0, len(splats), offset, subsampling=0, max_levels=9)
    check_build(orc.make_code, orig, res, size, offset)
    # radius replaced by 1/r^2 (kernels/octree.cl:193)
    np.testing.assert_array_equal(splats["radius"], np.float32(1.0) / (orig["radius"] * orig["radius"]))


def random_splats():
    # TestSplatTree::testRandom, test/test_splat_tree.cpp:172-200
    eng = MT19937()
    cells = (31, 31, 16)
    rows = []
    for _ in range(207):
        x = eng.uniform_real(-2.0, cells[0] + 2.0, np.float32)
        y = eng.uniform_real(-2.0, cells[1] + 2.0, np.float32)
        z = eng.uniform_real(-2.0, cells[2] + 2.0, np.float32)
        r = eng.uniform_real(0.25, 8.0, np.float32)
        rows.append((x, y, z, r))
    return make_splats(rows), cells


def test_random(orc):
    splats, cells = random_splats()
    sub = 2
    res = orc.octree_build(splats, 0, len(splats), (1, 2, -1), subsampling=sub, max_levels=8)
    start, commands = res["start"], res["commands"]
    for z in range(0, cells[2] + 1, 1 << sub):
        for y in range(0, cells[1] + 1, 1 << sub):
            for x in range(0, cells[0] + 1, 1 << sub):
                idx = orc.make_code(x >> sub, y >> sub, z >> sub)
                assert idx < len(start)
                found = walk_chain(start, commands, idx)
                assert all(0 <= c < len(splats) for c in found)


def test_empty_and_first_splat(orc):
    # zero splats: every start entry is -1
    res = orc.octree_build(make_splats([]), 0, 0, (0, 0, 0), subsampling=3, max_levels=6)
    assert res["used"] == 0 and (res["start"] == -1).all()
    # firstSplat offsets the ids stored in the command list
    rows = [(100.0, 100.0, 100.0, 1.0)] * 3 + [(4.5, 4.5, 4.5, 1.0)]
    splats = make_splats(rows)
    res = orc.octree_build(splats, 3, 1, (0, 0, 0), subsampling=3, max_levels=6)
    assert walk_chain(res["start"], res["commands"], 0) == [3]
    assert splats["radius"][0] == np.float32(1.0)   # untouched outside [first, first+n)
