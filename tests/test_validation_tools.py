"""plymanifold and meshdiff: the reference's acceptance tool (extras/plymanifold.cpp,
test/manifold.h) and a mesh comparison that ignores vertex numbering and triangle order."""
import os
import subprocess

import numpy as np
import pytest

from fixtures import is_manifold
from mlsgpu_b200 import plyfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "mlsgpu_b200", "host")


def build(name):
    subprocess.run(["make", "-C", os.path.join(ROOT, "mlsgpu_b200"), "host/" + name], check=True, capture_output=True)
    return os.path.join(HOST, name)


def write_mesh(path, verts, tris, ascii=False, index_type="uint32"):
    verts = np.asarray(verts, np.float32).reshape(-1, 3)
    tris = np.asarray(tris, np.int64).reshape(-1, 3)
    with open(path, "wb") as f:
        f.write(("ply\nformat %s 1.0\ncomment made by the tests\nelement vertex %d\nproperty float32 x\nproperty float32 y\n"
                 "property float32 z\nelement face %d\nproperty list uint8 %s vertex_indices\nend_header\n"
                 % ("ascii" if ascii else "binary_little_endian", len(verts), len(tris), index_type)).encode())
        if ascii:
            for v in verts:
                f.write(("%r %r %r\n" % tuple(float(x) for x in v)).encode())
            for t in tris:
                f.write(("3 %d %d %d\n" % tuple(t)).encode())
        else:
            f.write(verts.tobytes())
            rec = np.zeros(len(tris), np.dtype([("n", "u1"), ("i", "<" + {"uint32": "u4", "int32": "i4", "uint16": "u2"}[index_type], 3)]))
            rec["n"] = 3
            rec["i"] = tris
            f.write(rec.tobytes())


def plymanifold(tmp_path, verts, tris, **kw):
    exe = build("plymanifold")
    path = tmp_path / "m.ply"
    write_mesh(path, verts, tris, **kw)
    r = subprocess.run([exe, str(path)], capture_output=True, text=True)
    return r.returncode, r.stdout


def points(n):
    return [(i, i * i % 7, i * 3 % 5) for i in range(n)]


TETRA = [0, 1, 2, 0, 3, 1, 0, 2, 3, 1, 3, 2]


def report(out):
    lines = out.strip().splitlines()
    assert lines[0] == "Mesh is manifold."
    return {k: int(v) for k, v in (line.split(": ") for line in lines[1:])}


@pytest.mark.parametrize("kw", [{}, {"ascii": True}, {"index_type": "int32"}, {"index_type": "uint16"}])
def test_plymanifold_closed_and_open(tmp_path, kw):
    rc, out = plymanifold(tmp_path, points(4), TETRA, **kw)
    assert rc == 0 and report(out) == {"Vertices": 4, "Triangles": 4, "Components": 1, "Boundaries": 0}
    # two tetrahedra, one of them with a face removed: 2 components, 1 boundary loop
    tris = TETRA + [4 + i for i in TETRA[:9]]
    rc, out = plymanifold(tmp_path, points(8), tris, **kw)
    assert rc == 0 and report(out) == {"Vertices": 8, "Triangles": 7, "Components": 2, "Boundaries": 1}
    # a single triangle is one boundary
    rc, out = plymanifold(tmp_path, points(3), [0, 1, 2], **kw)
    assert rc == 0 and report(out)["Boundaries"] == 1
    # a vertex on two boundary loops counts them as one (test/manifold.h:85-91)
    rc, out = plymanifold(tmp_path, points(5), [0, 1, 2, 0, 3, 4], **kw)
    assert rc == 0 and report(out) == {"Vertices": 5, "Triangles": 2, "Components": 1, "Boundaries": 1}


@pytest.mark.parametrize("n,tris,reason", [
    (4, [0, 1, 4], "Triangle 0 contains out-of-range index 4"),
    (4, TETRA + [1, 1, 2], "Triangle 4 contains vertex 1 twice"),
    (5, TETRA, "Vertex 4 is isolated"),
    (4, [0, 1, 2, 0, 1, 3], "Edge 0 - 1 occurs twice with same winding"),
    (7, TETRA + [0, 4, 5, 0, 5, 6, 0, 6, 4], "Vertex 0 tunnels between interior regions"),
    (6, TETRA + [0, 4, 5], "Vertex 0 is both in the interior and on the boundary"),
])
def test_plymanifold_rejects(tmp_path, n, tris, reason):
    rc, out = plymanifold(tmp_path, points(n), tris)
    assert rc == 1 and out.strip() == "Mesh is not manifold: " + reason
    assert is_manifold(n, np.asarray(tris).reshape(-1, 3)) == reason    # the checker the other tests use agrees


def test_plymanifold_matches_python_checker_on_random_soups(tmp_path):
    rng = np.random.RandomState(5)
    verdicts = set()
    for case in range(150):
        n = int(rng.randint(3, 8))
        if case % 3 == 0:
            # perturb a closed surface: drop and flip a few faces of an octahedron
            octa = np.array([[0, 2, 4], [2, 1, 4], [1, 3, 4], [3, 0, 4], [2, 0, 5], [1, 2, 5], [3, 1, 5], [0, 3, 5]])
            keep = rng.rand(8) > 0.3
            tris = octa[keep]
            flip = rng.rand(len(tris)) > 0.8
            tris[flip] = tris[flip][:, ::-1]
            n = 6
        else:
            tris = rng.randint(0, n + (case % 7 == 0), size=(int(rng.randint(1, 9)), 3))
        if len(tris) == 0:
            continue
        expected = is_manifold(n, tris)
        rc, out = plymanifold(tmp_path, points(n), tris)
        if expected:
            assert rc == 1 and out.strip() == "Mesh is not manifold: " + expected, (n, tris.tolist())
        else:
            assert rc == 0 and out.startswith("Mesh is manifold."), (n, tris.tolist())
        verdicts.add(expected.split(" ")[0] if expected else "ok")
    assert {"ok", "Triangle", "Vertex", "Edge"} <= verdicts     # the soups reach every kind of verdict


def test_plymanifold_non_finite_and_format_errors(tmp_path):
    rc, out = plymanifold(tmp_path, [(0, 0, 0), (1, 0, 0), (0, np.nan, 0)], [0, 1, 2])
    assert rc == 1 and out.strip() == "Vertex 2 has non-finite value"
    exe = build("plymanifold")
    bad = tmp_path / "bad.ply"
    bad.write_bytes(b"ply\nformat binary_little_endian 1.0\nelement vertex 0\nproperty float32 x\nproperty float32 y\n"
                    b"element face 0\nproperty list uint8 uint32 vertex_indices\nend_header\n")
    r = subprocess.run([exe, str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "Missing property z" in r.stderr
    bad.write_bytes(b"ply\nformat binary_little_endian 1.0\nelement vertex 0\nproperty float32 x\nproperty float32 y\n"
                    b"property float32 z\nelement face 0\nproperty list uint8 float32 vertex_indices\nend_header\n")
    r = subprocess.run([exe, str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "should have integral type" in r.stderr
    quad = tmp_path / "quad.ply"
    quad.write_bytes(b"ply\nformat ascii 1.0\nelement vertex 4\nproperty float32 x\nproperty float32 y\nproperty float32 z\n"
                     b"element face 1\nproperty list uint8 uint32 vertex_indices\nend_header\n0 0 0\n1 0 0\n1 1 0\n0 1 0\n4 0 1 2 3\n")
    r = subprocess.run([exe, str(quad)], capture_output=True, text=True)
    assert r.returncode == 1 and "Face does not contain 3 vertices" in r.stderr
    assert subprocess.run([exe], capture_output=True).returncode == 2
    assert subprocess.run([exe, str(tmp_path / "missing.ply")], capture_output=True).returncode == 1


def test_plymanifold_reads_the_writers_output(tmp_path):
    """The tool accepts what PlyWriter produces (header padding comment included)."""
    exe = build("plymanifold")
    verts = np.asarray(points(4), np.float32)
    path = tmp_path / "w.ply"
    with open(path, "wb") as f:
        header = ("ply\nformat binary_little_endian 1.0\ncomment x\nelement vertex 4\nproperty float32 x\nproperty float32 y\n"
                  "property float32 z\nelement face 4\nproperty list uint8 uint32 vertex_indices\ncomment padding:XX\nend_header\n")
        f.write(header.encode())
        f.write(verts.tobytes())
        rec = np.zeros(4, np.dtype([("n", "u1"), ("i", "<u4", 3)]))
        rec["n"] = 3
        rec["i"] = np.asarray(TETRA).reshape(-1, 3)
        f.write(rec.tobytes())
    r = subprocess.run([exe, str(path)], capture_output=True, text=True)
    assert r.returncode == 0 and "Components: 1" in r.stdout
    v, t, _ = plyfile.read_mesh_ply(path)
    assert len(v) == 4 and len(t) == 4


def meshdiff(tmp_path, a, b, *args):
    exe = build("meshdiff")
    write_mesh(tmp_path / "a.ply", *a)
    write_mesh(tmp_path / "b.ply", *b)
    r = subprocess.run([exe] + list(args) + [str(tmp_path / "a.ply"), str(tmp_path / "b.ply")], capture_output=True, text=True)
    return r.returncode, r.stdout


def test_meshdiff_ignores_numbering_order_and_rotation(tmp_path):
    rng = np.random.RandomState(3)
    verts = rng.rand(30, 3).astype(np.float32)
    verts[0] = (0.0, -0.0, 1.0)
    tris = np.array([rng.choice(30, 3, replace=False) for _ in range(50)])
    perm = rng.permutation(30)
    inv = np.argsort(perm)
    verts2 = verts[perm]
    verts2[inv[0]] = (-0.0, 0.0, 1.0)                              # signed zeros compare equal
    tris2 = inv[tris]
    tris2 = np.array([np.roll(t, rng.randint(3)) for t in tris2])[rng.permutation(50)]
    rc, out = meshdiff(tmp_path, (verts, tris), (verts2, tris2))
    assert rc == 0 and "identical up to ordering" in out and "triangles: 50 common, 0 only in A, 0 only in B" in out
    # a flipped triangle is a different triangle
    tris3 = tris2.copy()
    tris3[7] = tris3[7][::-1]
    rc, out = meshdiff(tmp_path, (verts, tris), (verts2, tris3))
    assert rc == 1 and "1 only in A, 1 only in B" in out and "Meshes differ." in out
    # one moved vertex: exact mode sees it, a tolerance above the move does not
    verts4 = verts2.copy()
    verts4[3] += np.float32(1e-6)
    rc, out = meshdiff(tmp_path, (verts, tris), (verts4, tris2))
    assert rc == 1 and "vertices: 29 common, 1 only in A, 1 only in B" in out and "A-only vertex" in out
    rc, out = meshdiff(tmp_path, (verts, tris), (verts4, tris2), "--tolerance", "1e-3")
    assert rc == 0
    # an extra triangle and vertex on one side
    rc, out = meshdiff(tmp_path, (verts, tris), (np.vstack([verts2, [[9, 9, 9]]]), np.vstack([tris2, [[0, 1, 30]]])))
    assert rc == 1 and "0 only in A, 1 only in B" in out
    exe = build("meshdiff")
    assert subprocess.run([exe, "only-one.ply"], capture_output=True).returncode == 2
