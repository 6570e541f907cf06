"""numpy model of the lattice weld used by mlsgpu_b200/csrc/marching.cu.

The reference welds by sorting the per-cell vertices on their 64-bit keys and dropping duplicates
(kernels/marching.cl:184-340).  A vertex key is nothing but the position of the edge's midpoint on
the half-cell lattice (computeKey, :148-154: 2*cell + the two corner offsets), so the set of welded
vertices is the set of lattice points whose edge is crossed AND belongs to at least one valid
cell, their sorted order is the lattice's (z, y, x) order, and a vertex's welded index is a rank —
a prefix popcount — instead of the output of a sort.  This model states that in a few lines and
tests/test_lattice_model.py checks it against the oracle bit for bit; the CUDA kernels implement
the same thing with ballot masks.
"""
import numpy as np

# src/marching.cpp:50-81
EDGES = [(0, 1), (0, 2), (0, 3), (1, 3), (2, 3), (0, 4), (0, 5), (1, 5), (4, 5), (0, 6),
         (2, 6), (4, 6), (0, 7), (1, 7), (2, 7), (3, 7), (4, 7), (5, 7), (6, 7)]


def corner(k):
    return np.array([k & 1, (k >> 1) & 1, k >> 2])


def lattice_mesh(vol, tables, z_first, z_last, key_offset=(0, 0, 0)):
    """vol[z][y][x]; cells between corner slices z_first..z_last.  Returns the dict of
    Oracle.marching_generate: vertices (grid-global), keys, triangles, num_internal."""
    d, h, w = vol.shape
    cx, cy = w - 1, h - 1
    nz = z_last - z_first
    sign = vol >= 0
    finite = np.isfinite(vol)
    # cell validity and cube codes
    valid = np.ones((nz, cy, cx), bool)
    code = np.zeros((nz, cy, cx), np.int64)
    for k in range(8):
        ox, oy, oz = corner(k)
        sl = (slice(z_first + oz, z_first + oz + nz), slice(oy, oy + cy), slice(ox, ox + cx))
        valid &= finite[sl]
        code |= sign[sl].astype(np.int64) << k
    occupied = valid & (code != 0) & (code != 255)
    # lattice of edge midpoints: present <=> crossed by the surface in some valid cell
    present = np.zeros((2 * nz + 1, 2 * cy + 1, 2 * cx + 1), bool)
    for a, b in EDGES:
        oa, ob = corner(a), corner(b)
        dd = ob - oa
        assert (dd >= 0).all()
        crossed = valid & (((code >> a) & 1) != ((code >> b) & 1))
        lo = 2 * oa + dd          # lattice offset of the midpoint relative to 2 * cell
        present[lo[2]:lo[2] + 2 * nz:2, lo[1]:lo[1] + 2 * cy:2, lo[0]:lo[0] + 2 * cx:2] |= crossed
    z2, y2, x2 = np.nonzero(present)                       # (z, y, x) order = key order
    flagged = (x2 == 0) | (y2 == 0) | (z2 == 0) | (x2 == 2 * cx) | (y2 == 2 * cy)   # computeKey: top = (2cx, 2cy, 2 zFirst)
    external = flagged | (z2 == 2 * nz)                    # key >= minExternalKey = zLast << 43
    order = np.concatenate([np.nonzero(~flagged)[0], np.nonzero(flagged)[0]])
    rank = np.empty(len(order), np.int64)
    rank[order] = np.arange(len(order))
    num_internal = int((~external).sum())
    lookup = np.full(present.shape, -1, np.int64)
    lookup[z2, y2, x2] = rank
    # vertex positions: interp(), kernels/marching.cl:130-138, from the lower corner of the edge
    ax, ay, az = x2 >> 1, y2 >> 1, (z2 >> 1) + z_first
    dx, dy, dz = x2 & 1, y2 & 1, z2 & 1
    iso_a = vol[az, ay, ax].astype(np.float32)
    iso_b = vol[az + dz, ay + dy, ax + dx].astype(np.float32)
    with np.errstate(all="ignore"):
        inv = np.float32(1.0) / (iso_a - iso_b)
        t = (iso_a * inv).astype(np.float32)

    def fma(t, d, base):      # exact: t * d is t or 0
        return (t.astype(np.float64) * d + base.astype(np.float64)).astype(np.float32)

    ko = np.asarray(key_offset, np.int64)
    pos = np.stack([fma(t, dx, (ax + ko[0]).astype(np.float32)), fma(t, dy, (ay + ko[1]).astype(np.float32)),
                    fma(t, dz, (az + ko[2]).astype(np.float32))], axis=1)
    vertices = np.empty_like(pos)
    vertices[rank] = pos
    keys = np.zeros(len(rank), np.uint64)
    kz = z2 + 2 * z_first
    full = (x2.astype(np.uint64) | (y2.astype(np.uint64) << np.uint64(21)) | (kz.astype(np.uint64) << np.uint64(42)))
    full += np.uint64((int(ko[0]) << 1) | (int(ko[1]) << 22) | (int(ko[2]) << 43))
    keys[rank] = full
    # triangles: cells in x-fastest order, table order inside the cell
    count, start, data = tables["count"], tables["start"], tables["data"]
    tris = []
    for z, y, x in zip(*np.nonzero(occupied)):
        c = int(code[z, y, x])
        vbase, ibase = int(start[c][0]), int(start[c][1])
        for k in range(int(count[c][1])):
            e = int(data[vbase + int(data[ibase + k])])
            a, b = EDGES[e]
            lo = 2 * corner(a) + (corner(b) - corner(a))
            tris.append(lookup[2 * z + lo[2], 2 * y + lo[1], 2 * x + lo[0]])
    triangles = np.asarray(tris, np.uint32).reshape(-1, 3)
    return dict(vertices=vertices, keys=keys, triangles=triangles, num_internal=num_internal,
                external=np.sort(rank[external]))
