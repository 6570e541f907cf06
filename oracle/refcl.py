"""ctypes binding of oracle/_ref/librefcl.so: the REFERENCE'S OWN OpenCL kernels, run on the box's
OpenCL device by our C host (oracle/refcl/refcl.c).

ORACLE — TEST INFRASTRUCTURE ONLY (tests/ and bench.py's reference legs).  The library is built in
the build container, where /root/reference exists (`make -C oracle`), and travels to the GPU box
as a prebuilt file; `available()` says whether it is there and an OpenCL GPU answers.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "librefcl.so")
_P = C.c_void_p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


class RefCL:
    def __init__(self):
        if not os.path.exists(PATH):
            raise RuntimeError("%s missing: run `make -C oracle` where /root/reference is present" % PATH)
        lib = C.CDLL(PATH)
        lib.refcl_error.restype = C.c_char_p
        lib.refcl_device_name.restype = C.c_char_p
        lib.refcl_open.restype = C.c_int
        lib.refcl_octree_build.restype = C.c_int
        lib.refcl_octree_build.argtypes = [_P, C.c_uint32, C.c_uint32, C.c_uint32, _P, C.c_uint32, C.c_uint32, _P, _P, _P, _P, _P]
        lib.refcl_process_corners.restype = C.c_int
        lib.refcl_process_corners.argtypes = [_P, C.c_uint32, _P, C.c_uint32, _P, C.c_uint32, _P, C.c_uint32, C.c_uint32, _P,
                                              C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint32, C.c_uint32, C.c_float,
                                              C.c_int, C.c_uint32, C.c_uint32, _P, _P]
        lib.refcl_marching_generate.restype = C.c_int
        lib.refcl_marching_generate.argtypes = [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint32,
                                                C.c_uint32, _P, _P, _P, _P, _P, C.c_float, _P, _P, _P, _P, C.c_uint32,
                                                C.c_uint32, _P, _P]
        lib.refcl_solve_quadratic.restype = C.c_int
        lib.refcl_solve_quadratic.argtypes = [C.c_float, C.c_float, C.c_float, _P]
        self.lib = lib
        if lib.refcl_open() != 0:
            raise RuntimeError("refcl: " + lib.refcl_error().decode())
        self.device = lib.refcl_device_name().decode()
        self.seconds = 0.0          # accumulated kernel time (OpenCL profiling events)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("refcl: " + self.lib.refcl_error().decode())

    def octree_build(self, splats, first, n, offset, subsampling, max_levels, fill=-559038737):
        """SplatTreeCL::enqueueBuild with the reference kernels; mutates `splats`.  Same dict as Oracle."""
        max_shift = max_levels + subsampling - 1
        min_shift = min(subsampling, max_shift)
        num_start = sum(1 << (3 * (max_shift - i)) for i in range(min_shift, max_shift + 1))
        ne = 8 * n
        start = np.full(num_start, fill, np.int32)
        commands = np.full(ne + 2 * min(num_start, ne) + 1, fill, np.int32)
        keys = np.zeros(ne, np.uint32)
        values = np.zeros(ne, np.uint32)
        off = np.asarray(offset, np.int32)
        secs = C.c_double(0.0)
        self._check(self.lib.refcl_octree_build(_ptr(splats), len(splats), first, n, _ptr(off), subsampling, max_levels,
                                                _ptr(start), _ptr(commands), _ptr(keys), _ptr(values), C.byref(secs)))
        self.seconds += secs.value
        return dict(start=start, commands=commands, keys=keys, values=values, seconds=secs.value)

    def mls_process_corners(self, image, splats, commands, start, start_shift, offset, width, height, z_stride, z_bias,
                            z_first, z_last, boundary_factor, shape=0, max_swathe=0):
        """MlsFunctor::enqueue swathe by swathe.  Returns (kernel seconds, launches)."""
        assert image.dtype == np.float32 and image.flags.c_contiguous
        off = np.asarray(offset, np.int32)
        commands = np.ascontiguousarray(commands, np.int32)
        start = np.ascontiguousarray(start, np.int32)
        secs, launches = C.c_double(0.0), C.c_uint32(0)
        self._check(self.lib.refcl_process_corners(_ptr(image), image.shape[1], _ptr(splats), len(splats), _ptr(commands),
                                                   len(commands), _ptr(start), len(start), start_shift, _ptr(off), width,
                                                   height, z_stride, z_bias, z_first, z_last, boundary_factor, shape,
                                                   max_swathe, image.shape[0], C.byref(secs), C.byref(launches)))
        self.seconds += secs.value
        return secs.value, launches.value

    def marching_generate(self, image, width, height, z_stride, z_bias, z_first, z_last, key_offset, tables,
                          scale=1.0, bias=(0.0, 0.0, 0.0)):
        """One swathe + one shipOut with the reference kernels.  tables: Oracle.marching_tables() (the
        tables are host data in the reference, Marching::makeTables).  Same dict as Oracle."""
        assert image.dtype == np.float32 and image.flags.c_contiguous
        ncells = max(0, (width - 1) * (height - 1) * (z_last - z_first))
        vcap, icap = min(ncells * 13, 1 << 25) + 16, min(ncells * 36, 1 << 26) + 16
        verts = np.zeros((vcap, 3), np.float32)
        keys = np.zeros(vcap, np.uint64)
        idx = np.zeros(icap, np.uint32)
        sizes = np.zeros(5, np.uint32)
        key4 = np.zeros((2432, 4), np.uint32)
        key4[:, :3] = tables["key"]
        ko = np.asarray(key_offset, np.uint32)
        b = np.asarray(bias, np.float32)
        secs = C.c_double(0.0)
        count = np.ascontiguousarray(tables["count"], np.uint8)
        start = np.ascontiguousarray(tables["start"], np.uint16)
        data = np.ascontiguousarray(tables["data"], np.uint8)
        self._check(self.lib.refcl_marching_generate(_ptr(image), image.shape[1], width, height, z_stride, z_bias, z_first,
                                                     z_last, _ptr(ko), _ptr(count), _ptr(start), _ptr(data), _ptr(key4),
                                                     scale, _ptr(b), _ptr(verts), _ptr(keys), _ptr(idx), vcap, icap,
                                                     _ptr(sizes), C.byref(secs)))
        self.seconds += secs.value
        nv, ni, fe = int(sizes[0]), int(sizes[1]), int(sizes[2])
        return dict(vertices=verts[:nv].copy(), keys=keys[:nv].copy(), triangles=idx[:ni].reshape(-1, 3).copy(),
                    num_internal=fe, stats=(int(sizes[3]), int(sizes[4]), ni), seconds=secs.value)

    def solve_quadratic(self, a, b, c):
        out = np.zeros(1, np.float32)
        self._check(self.lib.refcl_solve_quadratic(a, b, c, _ptr(out)))
        return float(out[0])


_instance = None
_why = None


def get():
    """The shared RefCL instance, or None (with `why()` saying so) when the library or an OpenCL GPU
    is missing."""
    global _instance, _why
    if _instance is None and _why is None:
        try:
            _instance = RefCL()
        except (RuntimeError, OSError) as e:
            _why = str(e)
    return _instance


def why():
    return _why
