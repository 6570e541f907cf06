"""Pins the MLS oracle against the reference's known-answer tests (test/test_mls.cpp:255-514)."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from fixtures import sphere_splats, doubles_equal

EPS = float(np.finfo(np.float32).eps)


@pytest.fixture(scope="module", params=[False, True], ids=["fused", "as_written"])
def orc(request):
    return Oracle(as_written=request.param)


def test_make_code_decode(orc):
    # test/test_mls.cpp:255-285
    assert orc.make_code(2, 5, 3) == 174
    assert orc.decode(0xB1) == (1, 6, 2)
    assert orc.decode(0xAAAAAAAA) == (0x2AA, 0x555, 0x2AA)
    assert orc.decode(0x24924924) == (0, 0, 0x3FF)
    assert orc.decode(0) == (0, 0, 0)
    for lid in range(512):
        x, y, z = orc.decode(lid)
        assert 0 <= x < 8 and 0 <= y < 8 and 0 <= z < 8 and orc.make_code(x, y, z) == lid


def test_solve_quadratic(orc):
    # test/test_mls.cpp:287-330 (all 31 cases, 4 eps)
    n = float("nan")
    eps = 4 * EPS
    f = np.float32
    cases = [
        (n, -1, 2, -2), (n, -1e20, 2e10, -1.0001), (n, 1, 0, 1), (n, -1, 0, -1),
        (n, 0, 0, 0), (n, 0, 0, 4), (n, 0, 0, -3), (n, 0, 0, -1e20), (n, 0, 0, 1e20),
        (-1.5, 0, 2, 3), (0.0, 0, 5, 0), (0.0, 0, 1e20, 0), (0.0, 0, 1e-20, 0), (1e-20, 0, 1e10, 1e-10),
        (-1e20, 0, 1e-10, 1e10),
        (1.0, -1, 2, -1), (1.0, -10, 20, -10), (1e4, -1, 2e4, -1e8), (0.0, 1, 0, 0), (0.0, 1e30, 0, 0), (0.0, 1e-20, 0, 0),
        (2.0, -1, 5, -6), (2.0, -2, 10, -12), (2.0, 1, 1, -6), (2.0, f(0.1), f(0.1), f(-0.6)),
        (2.0, -1e-12, 5e-12, -6e-12), (-2e-12, 1, 5e-12, 6e-24),
        (1e-6, -1, 1 + 1e-6, -1e-6), (1.0, -1, 1 + 1e6, -1e6), (1e20, -1e-20, 2, -1e20), (-1e-6, 1e-6, 1, 1e-6),
    ]
    assert len(cases) == 31
    for want, a, b, c in cases:
        got = orc.solve_quadratic(f(a), f(b), f(c))
        assert doubles_equal(want, got, eps), (want, got, a, b, c)


def make_sphere(xc, yc, zc, r, grad):
    # test/test_mls.cpp:154-168
    f = np.float32
    scale = f(grad) * f(0.5) / f(r)
    return [f(-2.0) * f(xc) * scale, f(-2.0) * f(yc) * scale, f(-2.0) * f(zc) * scale, scale,
            (f(xc) * f(xc) + f(yc) * f(yc) + f(zc) * f(zc) - f(r) * f(r)) * scale]


def make_plane(px, py, pz, dx, dy, dz):
    return [dx, dy, dz, 0.0, -(dx * px + dy * py + dz * pz)]


def test_project_dist_origin_sphere(orc):
    # test/test_mls.cpp:332-347
    eps = 4 * EPS
    cases = [(7.0, make_sphere(3, 4, 12, 6, 1.0)), (7.0, make_sphere(3, 4, 12, 6, 2.5)),
             (-7.0, make_sphere(3, 4, 12, 6, -2.5)), (0.0, make_sphere(3, 4, 12, 13, 2.5)),
             (-5.0, make_sphere(3, 4, 12, 18, 2.5)), (-6.0, make_sphere(0, 0, 0, 6, 2.5)),
             (5.0, make_sphere(0, 0, 0, 5, -1.5)),
             (-5.0 / 1.5, make_plane(1, 2, 3, 1, 0.5, 1)), (5.0 / 1.5, make_plane(-1, -2, -3, 1, 0.5, 1))]
    for want, params in cases:
        got = orc.project_dist_origin_sphere(params)
        assert doubles_equal(want, got, eps), (want, got)


def test_fit_sphere(orc):
    # test/test_mls.cpp:382-409: all inputs lie on the fitted sphere, gradient = normal, 16 eps
    eps = 16 * EPS
    splats = sphere_splats(20, (1.0, 2.0, 3.5), 6.5)
    params = orc.fit_sphere(splats).astype(np.float64)
    for s in splats:
        x, y, z = (float(v) for v in s["position"])
        v = params[0] * x + params[1] * y + params[2] * z + params[3] * (x * x + y * y + z * z) + params[4]
        assert doubles_equal(0.0, v, eps)
        g = [np.float32(2 * params[3] * c + params[i]) for i, c in enumerate((x, y, z))]
        for i in range(3):
            assert doubles_equal(float(s["normal"][i]), float(g[i]), eps)


def process_corners_case(orc, radius=65.0):
    """TestMls::testProcessCorners inputs, test/test_mls.cpp:416-480.

    With the upstream radius (65) every corner of the region lies more than sqrt(3) inside the
    sphere, so the upstream assertions only ever see NaN; radius=35 puts the surface through the
    region and makes the same assertions compare real distances."""
    N = 50
    center = (10.0, 20.0, 35.0)
    size = (19, 24, 28)
    offset = (20, 15, 33)
    wgs = 8
    image_w = -(-size[0] // wgs) * wgs
    image_h = -(-size[1] // wgs) * wgs
    image_d = -(-size[2] // wgs) * wgs
    z_first, z_last = wgs, 26
    z_stride = image_h + 10
    z_bias = (2 - z_first) * z_stride
    sub = 3
    while (2 << sub) < max(size):
        sub += 1
    splats = sphere_splats(N, center, radius)
    splats["radius"] = np.float32(1.0) / (splats["radius"] * splats["radius"])
    start = np.zeros(8, np.int32)
    commands = [N - 1] + list(range(N - 2))
    commands.append(-2 - len(commands))
    start[6] = -1
    start[7] = len(commands)
    commands.append(len(commands) + 3)
    commands += [N - 2, N - 1, -1]
    commands = np.array(commands, np.int32)
    image = np.full((image_d * z_stride + z_bias, image_w), 12345.0, np.float32)
    return dict(splats=splats, start=start, commands=commands, image=image, sub=sub, offset=offset,
                size=size, z_first=z_first, z_last=z_last, z_stride=z_stride, z_bias=z_bias,
                center=center, radius=radius)


def check_process_corners(c, image):
    size, offset, sub = c["size"], c["offset"], c["sub"]
    cx, cy, cz = c["center"]
    nan_count = 0
    for z in range(c["z_first"], c["z_last"] + 1):
        for y in range(size[1]):
            for x in range(size[0]):
                px, py, pz = np.float32(x + offset[0]), np.float32(y + offset[1]), np.float32(z + offset[2])
                expected = np.float32(np.sqrt((float(px) - cx) ** 2 + (float(py) - cy) ** 2 + (float(pz) - cz) ** 2)) - np.float32(c["radius"])
                if (z >> sub) == 1 and (y >> sub) == 1:
                    expected = np.nan
                if abs(expected) > np.sqrt(np.float32(3.0)):
                    expected = np.nan
                actual = image[y + z * c["z_stride"] + c["z_bias"], x]
                nan_count += int(np.isnan(actual))
                assert doubles_equal(float(expected), float(actual), 1e-5), (x, y, z, expected, actual)
    assert 0 < nan_count
    return nan_count


@pytest.mark.parametrize("radius", [65.0, 35.0])
def test_process_corners(orc, radius):
    c = process_corners_case(orc, radius)
    t, h = orc.mls_process_corners(c["image"], c["splats"], c["commands"], c["start"], 3 * c["sub"],
                                   c["offset"], c["size"][0], c["size"][1], c["z_stride"], c["z_bias"],
                                   c["z_first"], c["z_last"], orc.boundary_factor(1.0), shape=0)
    assert t > 0 and 0 < h <= t
    nan_count = check_process_corners(c, c["image"])
    total = c["size"][0] * c["size"][1] * (c["z_last"] - c["z_first"] + 1)
    if radius == 35.0:
        assert total - nan_count > 300     # real distances were compared


def test_fused_matches_as_written():
    """The two legal roundings of the accumulation agree far inside the 1e-5 tolerance."""
    a, b = Oracle(False), Oracle(True)
    ca, cb = process_corners_case(a), process_corners_case(b)
    for o, c in ((a, ca), (b, cb)):
        o.mls_process_corners(c["image"], c["splats"], c["commands"], c["start"], 3 * c["sub"], c["offset"],
                              c["size"][0], c["size"][1], c["z_stride"], c["z_bias"], c["z_first"],
                              c["z_last"], o.boundary_factor(1.0), shape=0)
    ia, ib = ca["image"], cb["image"]
    assert (np.isnan(ia) == np.isnan(ib)).all()
    m = ~np.isnan(ia)
    assert np.abs(ia[m] - ib[m]).max() <= 1e-5


def test_boundary_factor(orc):
    g = np.sqrt(6.0) * 512 / (693 * np.pi)
    assert orc.boundary_factor(1.0) == pytest.approx(1 - g * g, rel=1e-5)
    assert orc.boundary_factor(0.0) == 1.0
