"""Pins the marching oracle against the reference's known-answer tests
(test/test_marching.cpp:270-632, test/test_mesh_filter.cpp:284-361)."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from fixtures import is_manifold

EDGES = [(0, 1), (0, 2), (0, 3), (1, 3), (2, 3), (0, 4), (0, 5), (1, 5), (4, 5), (0, 6),
         (2, 6), (4, 6), (0, 7), (1, 7), (2, 7), (3, 7), (4, 7), (5, 7), (6, 7)]


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def test_tables(orc):
    # TestMarching::testConstructor, test/test_marching.cpp:270-305, + sizes src/marching.cpp:248-251
    t = orc.marching_tables()
    count, start, data = t["count"], t["start"], t["data"]
    assert t["sizes"] == (2432, 8192)
    assert int(start[-1][1]) == 8192 and int(start[0][0]) == 0
    for i in range(256):
        sv, si = int(start[i][0]), int(start[i][1])
        ev, ei = int(start[i + 1][0]), int(start[i + 1][1])
        assert int(count[i][0]) == ev - sv and int(count[i][1]) == ei - si
        assert count[i][1] % 3 == 0
        assert count[i][0] <= 13 and count[i][1] <= 36
        verts = data[sv:ev]
        assert (np.diff(verts.astype(int)) > 0).all() and (verts < 19).all()
        assert (data[si:ei] < ev - sv).all()
        # every emitted vertex lies on an edge whose end points have different signs, and
        # every sign-change edge is emitted (what makes vertex welding by key possible)
        want = [e for e, (a, b) in enumerate(EDGES) if ((i >> a) & 1) != ((i >> b) & 1)]
        assert list(verts) == want
        for j, e in enumerate(verts):
            a, b = EDGES[e]
            for axis in range(3):
                assert t["key"][sv + j][axis] == ((a >> axis) & 1) + ((b >> axis) & 1)
    assert count[0].tolist() == [0, 0] and count[255].tolist() == [0, 0]


def make_key(x, y, z, external):
    return (z << 42) | (y << 21) | x | ((1 << 63) if external else 0)


def test_compute_key(orc):
    # test/test_marching.cpp:333-351 (all 11 cases)
    cases = [((0, 0, 0), (32, 32, 32), True), ((1, 2, 3), (32, 32, 32), False), ((0, 4, 5), (32, 32, 32), True),
             ((6, 0, 7), (32, 32, 32), True), ((9, 5, 0), (32, 32, 32), False), ((30, 1, 2), (30, 40, 50), True),
             ((5, 40, 3), (30, 40, 50), True), ((1, 2, 50), (30, 40, 50), True), ((1, 2, 40), (30, 40, 50), False),
             ((1, 2, 30), (30, 40, 50), False), ((30, 40, 50), (30, 40, 50), True)]
    for c, top, ext in cases:
        assert orc.compute_key(c, top) == make_key(*c, ext)


def test_compact_vertices(orc):
    # test/test_marching.cpp:401-479: mixed / all external / all internal
    ext = 1 << 63
    in_keys = np.array([100, 100, 200, ext | 50, ext | 50, 0xFFFFFFFFFFFFFFFF], np.uint64)
    unique = np.array([0, 0, 1, 2, 2, 3], np.uint32)
    ids = [4, 1, 2, 3, 0]
    in_verts = np.zeros((5, 4), np.float32)
    for i in range(5):
        in_verts[i, :3] = (i, i + 1, i + 2)
        in_verts[i, 3] = np.array([ids[i]], np.uint32).view(np.float32)[0]
    dead = 0xDEADBEEFDEADBEEF
    for min_ext, want_keys, want_first in ((200, [dead, 200, 50], 1), (100, [100, 200, 50], 0),
                                           (ext | 60, [dead, dead, dead], 3)):
        out_v = np.zeros((3, 3), np.float32)
        out_k = np.full(3, dead, np.uint64)
        remap = np.full(5, 0xDEADBEEF, np.uint32)
        first = orc.compact_vertices(out_v, out_k, remap, unique, in_verts, in_keys, min_ext, 0)
        assert out_v[:, 0].tolist() == [1.0, 2.0, 4.0]
        assert out_k.tolist() == want_keys
        assert remap.tolist() == [2, 0, 1, 2, 0]
        assert first == want_first


def test_scale_bias(orc):
    # test/test_mesh_filter.cpp:284-361 (5 vertices x (3, (10,-20,30)), 1e-2)
    v = np.array([[0, 0, 0], [1, 2, 3], [-1, 5, 0.5], [100, -100, 7], [0.25, 0.5, 0.75]], np.float32)
    want = v.astype(np.float64) * 3 + np.array([10, -20, 30])
    orc.scale_bias(v, 3.0, (10.0, -20.0, 30.0))
    np.testing.assert_allclose(v, want, atol=1e-2)
    orc.scale_bias(np.zeros((0, 3), np.float32), 3.0, (0, 0, 0))  # empty mesh is fine


def sphere_field(w, h, d, cx, cy, cz, radius):
    # SphereGenerator::generate, test/test_marching.cpp:143-147 (including its (y - cx) quirk)
    f = np.float32
    z, y, x = np.meshgrid(np.arange(d, dtype=f), np.arange(h, dtype=f), np.arange(w, dtype=f), indexing="ij")
    d2 = (x - f(cx)) * (x - f(cx)) + (y - f(cx)) * (y - f(cy)) + (z - f(cz)) * (z - f(cz))
    with np.errstate(invalid="ignore"):   # the quirk makes d2 negative in places: NaN, as upstream
        return (np.sqrt(d2.astype(f)) - f(radius)).astype(f)


def alternating_field(w, h, d):
    z, y, x = np.meshgrid(np.arange(d), np.arange(h), np.arange(w), indexing="ij")
    return np.where(((x ^ y ^ z) & 1) == 1, 1.0, -1.0).astype(np.float32)


def to_image(vol, z_stride=None, pad_w=0):
    """[D][H][W] volume -> 2-D image with rows y + z*zStride (zBias = 0)."""
    d, h, w = vol.shape
    z_stride = z_stride or h
    img = np.full((d * z_stride, w + pad_w), np.nan, np.float32)
    for z in range(d):
        img[z * z_stride:z * z_stride + h, :w] = vol[z]
    return img, z_stride


def weld_by_key(blocks):
    """Host-mesher stand-in: concatenates shipped-out blocks, merging external vertices that
    share a key (src/mesher.cpp:280-306)."""
    verts, tris, keymap = [], [], {}
    for b in blocks:
        ni = b["num_internal"]
        remap = np.zeros(len(b["vertices"]), np.int64)
        for i, v in enumerate(b["vertices"]):
            if i >= ni:
                k = int(b["keys"][i])
                if k in keymap:
                    remap[i] = keymap[k]
                    continue
                keymap[k] = len(verts)
            remap[i] = len(verts)
            verts.append(v)
        tris.extend(remap[b["triangles"]].tolist())
    return np.array(verts), np.array(tris).reshape(-1, 3)


MARCHING_CASES = {
    # test/test_marching.cpp:594-632
    "sphere": lambda: sphere_field(71, 75, 60, 30.0, 41.5, 27.75, 25.3),
    "truncated": lambda: sphere_field(71, 75, 60, 0.5 * 71, 0.5 * 75, 0.5 * 60, 42.0),
    "alternating": lambda: alternating_field(32, 32, 32),
}


@pytest.mark.parametrize("name", sorted(MARCHING_CASES))
def test_generate_manifold(orc, name):
    vol = MARCHING_CASES[name]()
    d, h, w = vol.shape
    img, zs = to_image(vol, z_stride=h + 3, pad_w=5)
    res = orc.marching_generate(img, w, h, zs, 0, 0, d - 1, (0, 0, 0))
    nv = len(res["vertices"])
    assert nv > 0 and len(res["triangles"]) > 0
    assert is_manifold(nv, res["triangles"]) == ""
    # internal vertices first, external after; external = on the faces of the block
    ni = res["num_internal"]
    v = res["vertices"]
    on_face = ((v[:, 0] == 0) | (v[:, 1] == 0) | (v[:, 2] == 0) | (v[:, 0] == w - 1) | (v[:, 1] == h - 1)
               | (v[:, 2] == d - 1))
    assert not on_face[:ni].any() and on_face[ni:].all()
    # split into two shipped-out blocks and weld by key: same mesh size, still manifold
    zmid = 11 if name != "alternating" else 16
    b0 = orc.marching_generate(img, w, h, zs, 0, 0, zmid, (0, 0, 0))
    b1 = orc.marching_generate(img, w, h, zs, 0, zmid, d - 1, (0, 0, 0))
    verts, tris = weld_by_key([b0, b1])
    assert len(verts) == nv and len(tris) == len(res["triangles"])
    assert is_manifold(len(verts), tris) == ""


def test_generate_empty_and_nan(orc):
    vol = np.full((4, 5, 6), 1.0, np.float32)
    img, zs = to_image(vol)
    res = orc.marching_generate(img, 6, 5, zs, 0, 0, 3, (0, 0, 0))
    assert len(res["vertices"]) == 0 and len(res["triangles"]) == 0
    # a NaN corner kills the (up to 8) cells around it
    vol = sphere_field(12, 12, 12, 5.5, 5.5, 5.5, 3.2)
    vol[:, :, :] = sphere_field(12, 12, 12, 5.5, 5.5, 5.5, 3.2)
    img, zs = to_image(vol)
    full = orc.marching_generate(img, 12, 12, zs, 0, 0, 11, (0, 0, 0))
    vol[5, 2, 5] = np.nan
    img, zs = to_image(vol)
    holed = orc.marching_generate(img, 12, 12, zs, 0, 0, 11, (0, 0, 0))
    assert 0 < len(holed["triangles"]) < len(full["triangles"])


def test_key_offset(orc):
    vol = sphere_field(12, 12, 12, 5.5, 5.5, 5.5, 7.0)   # truncated by the box: has external vertices
    img, zs = to_image(vol)
    a = orc.marching_generate(img, 12, 12, zs, 0, 0, 11, (0, 0, 0))
    b = orc.marching_generate(img, 12, 12, zs, 0, 0, 11, (7, 100, 3))
    ni = a["num_internal"]
    assert ni == b["num_internal"] and len(a["vertices"]) > ni
    # positions are interpolated in grid-global coordinates, so only equal up to rounding
    np.testing.assert_allclose(b["vertices"], a["vertices"] + np.array([7, 100, 3], np.float32), atol=2e-5)
    off = (7 << 1) | (100 << 22) | (3 << 43)
    np.testing.assert_array_equal(b["keys"][ni:], a["keys"][ni:] + np.uint64(off))
    np.testing.assert_array_equal(a["triangles"], b["triangles"])
