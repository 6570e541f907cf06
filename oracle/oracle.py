"""ctypes binding of the CPU oracle (oracle/liboracle.so).

ORACLE — TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package (mlsgpu_b200/) must
never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

SPLAT_DTYPE = np.dtype([("position", np.float32, 3), ("radius", np.float32),
                        ("normal", np.float32, 3), ("quality", np.float32)])
assert SPLAT_DTYPE.itemsize == 32


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def _load(name):
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    try:
        return C.CDLL(path)
    except OSError:
        build()
        return C.CDLL(path)


def _ptr(a, typ=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, as_written=False):
        lib = _load("liboracle_aswritten.so" if as_written else "liboracle.so")
        self.lib = lib
        lib.orc_level_shift.restype = C.c_int
        lib.orc_level_shift.argtypes = [C.c_int] * 6
        lib.orc_point_box_dist2.restype = C.c_float
        lib.orc_point_box_dist2.argtypes = [C.c_void_p] * 3
        lib.orc_make_code.restype = C.c_uint32
        lib.orc_make_code.argtypes = [C.c_int] * 3
        lib.orc_decode.restype = None
        lib.orc_decode.argtypes = [C.c_uint32, C.c_void_p]
        lib.orc_solve_quadratic.restype = C.c_float
        lib.orc_solve_quadratic.argtypes = [C.c_float] * 3
        lib.orc_project_dist_origin_sphere.restype = C.c_float
        lib.orc_project_dist_origin_sphere.argtypes = [C.c_void_p]
        lib.orc_fit_sphere.restype = None
        lib.orc_fit_sphere.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        lib.orc_boundary_factor.restype = C.c_float
        lib.orc_boundary_factor.argtypes = [C.c_float]
        lib.orc_octree_num_start.restype = C.c_uint32
        lib.orc_octree_num_start.argtypes = [C.c_uint32, C.c_uint32]
        lib.orc_write_entries.restype = None
        lib.orc_write_entries.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        lib.orc_octree_build.restype = C.c_uint32
        lib.orc_octree_build.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32,
                                         C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]
        lib.orc_mls_process_corners.restype = None
        lib.orc_mls_process_corners.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                                C.c_int32, C.c_uint32, C.c_uint32, C.c_float, C.c_int,
                                                C.c_void_p]
        lib.orc_marching_tables.restype = None
        lib.orc_marching_tables.argtypes = [C.c_void_p] * 5
        lib.orc_compute_key.restype = C.c_uint64
        lib.orc_compute_key.argtypes = [C.c_void_p, C.c_void_p]
        lib.orc_compact_vertices.restype = None
        lib.orc_compact_vertices.argtypes = [C.c_void_p] * 7 + [C.c_uint64, C.c_uint64, C.c_uint32]
        lib.orc_scale_bias.restype = None
        lib.orc_scale_bias.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_void_p]
        lib.orc_marching_generate.restype = C.c_int
        lib.orc_marching_generate.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                              C.c_int32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                              C.c_void_p]
        lib.orc_condition_splat.restype = None
        lib.orc_condition_splat.argtypes = [C.c_void_p, C.c_float, C.c_float]
        lib.orc_world_to_grid.restype = None
        lib.orc_world_to_grid.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_float, C.c_void_p]

    # ---- scalar helpers ------------------------------------------------------------------
    def level_shift(self, lo, hi):
        return self.lib.orc_level_shift(*[int(v) for v in lo], *[int(v) for v in hi])

    def point_box_dist2(self, pos, lo, hi):
        a = [np.asarray(v, np.float32) for v in (pos, lo, hi)]
        return float(self.lib.orc_point_box_dist2(*[_ptr(v) for v in a]))

    def make_code(self, x, y, z):
        return int(self.lib.orc_make_code(int(x), int(y), int(z)))

    def decode(self, code):
        out = np.zeros(3, np.int32)
        self.lib.orc_decode(int(code) & 0xFFFFFFFF, _ptr(out))
        return tuple(int(v) for v in out)

    def solve_quadratic(self, a, b, c):
        return float(self.lib.orc_solve_quadratic(a, b, c))

    def project_dist_origin_sphere(self, params):
        p = np.asarray(params, np.float32)
        return float(self.lib.orc_project_dist_origin_sphere(_ptr(p)))

    def fit_sphere(self, splats):
        out = np.zeros(5, np.float32)
        self.lib.orc_fit_sphere(_ptr(splats), len(splats), _ptr(out))
        return out

    def boundary_factor(self, limit):
        return float(self.lib.orc_boundary_factor(limit))

    # ---- octree --------------------------------------------------------------------------
    def num_start(self, max_levels, subsampling):
        return int(self.lib.orc_octree_num_start(max_levels, subsampling))

    def octree_build(self, splats, first, n, offset, subsampling, max_levels, fill=-559038737):
        """Returns dict(start, commands, used, keys, values, commandMap); mutates `splats`."""
        num_start = self.num_start(max_levels, subsampling)
        ne = 8 * n
        start = np.full(num_start, fill, np.int32)
        commands = np.full(ne + 2 * min(num_start, ne) + 1, fill, np.int32)
        keys = np.zeros(ne, np.uint32)
        values = np.zeros(ne, np.uint32)
        cmap = np.zeros(ne, np.uint32)
        off = np.asarray(offset, np.int32)
        used = self.lib.orc_octree_build(_ptr(splats), first, n, _ptr(off), subsampling, max_levels,
                                         _ptr(start), _ptr(commands), _ptr(keys), _ptr(values), _ptr(cmap))
        return dict(start=start, commands=commands, used=int(used), keys=keys, values=values,
                    commandMap=cmap)

    # ---- MLS -----------------------------------------------------------------------------
    def mls_process_corners(self, image, splats, commands, start, start_shift, offset, width, height,
                            z_stride, z_bias, z_first, z_last, boundary_factor, shape=0):
        """image: 2-D float32 array (rows = y + z*zStride + zBias, pitch = image.shape[1]).
        Returns (pair tests, hits)."""
        assert image.dtype == np.float32 and image.flags.c_contiguous
        off = np.asarray(offset, np.int32)
        stats = np.zeros(2, np.uint64)
        commands = np.ascontiguousarray(commands, np.int32)
        start = np.ascontiguousarray(start, np.int32)
        self.lib.orc_mls_process_corners(_ptr(image), image.shape[1], _ptr(splats), _ptr(commands),
                                         _ptr(start), start_shift, _ptr(off), width, height, z_stride,
                                         z_bias, z_first, z_last, boundary_factor, shape, _ptr(stats))
        return int(stats[0]), int(stats[1])

    # ---- marching ------------------------------------------------------------------------
    def marching_tables(self):
        count = np.zeros((256, 2), np.uint8)
        start = np.zeros((257, 2), np.uint16)
        data = np.zeros(8192, np.uint8)
        key = np.zeros((2432, 3), np.uint32)
        sizes = np.zeros(2, np.uint32)
        self.lib.orc_marching_tables(_ptr(count), _ptr(start), _ptr(data), _ptr(key), _ptr(sizes))
        return dict(count=count, start=start, data=data, key=key, sizes=(int(sizes[0]), int(sizes[1])))

    def compute_key(self, coords, top):
        c = np.asarray(coords, np.uint32)
        t = np.asarray(top, np.uint32)
        return int(self.lib.orc_compute_key(_ptr(c), _ptr(t)))

    def compact_vertices(self, out_vertices, out_keys, index_remap, vertex_unique, in_vertices4, in_keys,
                         min_external_key, key_offset):
        first_external = np.full(1, 0xDEADBEEF, np.uint32)
        n = len(in_vertices4)
        self.lib.orc_compact_vertices(_ptr(out_vertices), _ptr(out_keys), _ptr(index_remap),
                                      _ptr(first_external), _ptr(vertex_unique), _ptr(in_vertices4),
                                      _ptr(in_keys), min_external_key, key_offset, n)
        return int(first_external[0])

    def scale_bias(self, vertices, scale, bias):
        b = np.asarray(bias, np.float32)
        self.lib.orc_scale_bias(_ptr(vertices), len(vertices), scale, _ptr(b))

    def marching_generate(self, image, width, height, z_stride, z_bias, z_first, z_last, key_offset,
                          vcap=None, icap=None):
        """Returns dict(vertices (nV,3), keys (nV,), triangles (nT,3), num_internal, stats)."""
        assert image.dtype == np.float32 and image.flags.c_contiguous
        ncells = max(0, (width - 1) * (height - 1) * (z_last - z_first))
        if vcap is None:
            vcap = min(ncells * 13, 1 << 26) + 16
        if icap is None:
            icap = min(ncells * 36, 1 << 27) + 16
        verts = np.zeros((vcap, 3), np.float32)
        keys = np.full(vcap, 0xDEADBEEFDEADBEEF, np.uint64)
        idx = np.zeros(icap, np.uint32)
        sizes = np.zeros(3, np.uint32)
        stats = np.zeros(3, np.uint32)
        ko = np.asarray(key_offset, np.uint32)
        rc = self.lib.orc_marching_generate(_ptr(image), image.shape[1], width, height, z_stride, z_bias,
                                            z_first, z_last, _ptr(ko), _ptr(verts), _ptr(keys), _ptr(idx),
                                            vcap, icap, _ptr(sizes), _ptr(stats))
        if rc != 0:
            raise MemoryError("oracle marching capacity exceeded")
        nv, nt, ni = (int(v) for v in sizes)
        return dict(vertices=verts[:nv].copy(), keys=keys[:nv].copy(),
                    triangles=idx[:3 * nt].reshape(nt, 3).copy(), num_internal=ni,
                    stats=tuple(int(v) for v in stats))

    # ---- host conditioning ---------------------------------------------------------------
    def condition_splats(self, splats, smooth, max_radius=np.inf):
        """FastPly::Reader::decode semantics, vectorised the same way (double division)."""
        r = np.minimum(splats["radius"], np.float32(max_radius)) * np.float32(smooth)
        splats["radius"] = r
        splats["quality"] = (1.0 / (r * r).astype(np.float64)).astype(np.float32)

    def world_to_grid(self, splats, reference, spacing, lo):
        ref = np.asarray(reference, np.float32)
        lo = np.asarray(lo, np.int32)
        self.lib.orc_world_to_grid(_ptr(splats), len(splats), _ptr(ref), spacing, _ptr(lo))
