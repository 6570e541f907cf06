"""Binary PLY helpers for tests and tools: splat input files and mesh output files in the formats
of src/fast_ply.cpp (reader :185-350, writer :443-564)."""
import numpy as np


def write_splat_ply(path, splats, extra_property=False):
    """splats: SPLAT_DTYPE array in WORLD units (radius = raw sample radius)."""
    n = len(splats)
    fields = [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4")]
    if extra_property:
        fields.append(("confidence", "u1"))      # an unrelated property the reader must skip
    fields.append(("radius", "<f4"))
    rec = np.zeros(n, np.dtype(fields))
    for i, k in enumerate(("x", "y", "z")):
        rec[k] = splats["position"][:, i]
    for i, k in enumerate(("nx", "ny", "nz")):
        rec[k] = splats["normal"][:, i]
    rec["radius"] = splats["radius"]
    names = {"<f4": "float32", "u1": "uint8"}
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % n
    header += "".join("property %s %s\n" % (names[t], k) for k, t in fields) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode())
        f.write(rec.tobytes())


def read_mesh_ply(path):
    """Returns (vertices float32 (n,3), triangles uint32 (m,3), header text)."""
    data = open(path, "rb").read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    header = data[:end].decode()
    nv = nt = 0
    for line in header.splitlines():
        if line.startswith("element vertex"):
            nv = int(line.split()[2])
        if line.startswith("element face"):
            nt = int(line.split()[2])
    verts = np.frombuffer(data, "<f4", nv * 3, end).reshape(nv, 3)
    faces = np.frombuffer(data, np.uint8, nt * 13, end + nv * 12).reshape(nt, 13)
    assert (faces[:, 0] == 3).all()
    tris = np.ascontiguousarray(faces[:, 1:]).view("<u4").reshape(nt, 3)
    assert len(data) == end + nv * 12 + nt * 13
    return verts, tris, header
