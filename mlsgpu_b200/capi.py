"""ctypes binding of include/mlsb200.h.

Method names follow the reference's classes: ``octree_build`` is SplatTreeCL::enqueueBuild,
``mls_set`` / ``mls_corners`` are MlsFunctor::set / enqueue, ``marching_generate`` is
Marching::generate (file:line citations are in the header).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

SPLAT_DTYPE = np.dtype([("position", np.float32, 3), ("radius", np.float32),
                        ("normal", np.float32, 3), ("quality", np.float32)])

STATUS_NAMES = {0: "MLSB_OK", 1: "MLSB_ERR_INVALID_ARGUMENT", 2: "MLSB_ERR_LENGTH", 3: "MLSB_ERR_STATE",
                4: "MLSB_ERR_CUDA", 5: "MLSB_ERR_OVERFLOW"}


class MlsbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, status), message))
        self.status = status


class WorkerConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("maxLevels", C.c_uint32), ("subsampling", C.c_uint32),
                ("maxSplats", C.c_uint32), ("maxCells", C.c_uint32), ("meshBytes", C.c_uint64),
                ("shape", C.c_int32), ("boundaryLimit", C.c_float)]


class Swathe(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("zStride", C.c_uint32), ("zBias", C.c_int32),
                ("zFirst", C.c_uint32), ("zLast", C.c_uint32)]


class MeshSizes(C.Structure):
    _fields_ = [("numVertices", C.c_uint64), ("numTriangles", C.c_uint64), ("numInternalVertices", C.c_uint64)]


class ResourceUsage(C.Structure):
    _fields_ = [("deviceBytes", C.c_uint64), ("maxAllocationBytes", C.c_uint64), ("pinnedHostBytes", C.c_uint64)]


class DeviceMesh(C.Structure):
    _fields_ = [("sizes", MeshSizes), ("vertices", C.c_void_p), ("vertexKeys", C.c_void_p), ("triangles", C.c_void_p)]


MAX_LEVELS, MAX_SPLATS, MAX_DIMENSION, MAX_GLOBAL_DIMENSION = 10, (1 << 31) // 16, 512, (1 << 20) - 1

MESH_SINK = C.CFUNCTYPE(None, C.POINTER(MeshSizes), C.c_void_p, C.c_void_p)
FIELD_GENERATOR = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(Swathe), C.c_void_p)
DEVICE_MESH_SINK = C.CFUNCTYPE(C.c_int, C.POINTER(DeviceMesh), C.c_void_p, C.c_void_p)
MESH_FILTER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(DeviceMesh), C.POINTER(DeviceMesh), C.c_void_p)
GROUP_SINK = C.CFUNCTYPE(None, C.c_uint64, C.POINTER(MeshSizes), C.c_void_p, C.c_void_p)

# every symbol include/mlsb200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
API = {
    "mlsb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "mlsb_last_error_string": (C.c_char_p, []),
    "mlsb_version": (C.c_char_p, []),
    "mlsb_worker_create": (C.c_int, [C.POINTER(WorkerConfig), C.POINTER(_P)]),
    "mlsb_worker_destroy": (None, [_P]),
    "mlsb_worker_resource_usage": (C.c_int, [C.POINTER(WorkerConfig), C.POINTER(C.c_uint64)]),
    "mlsb_worker_resource_usage_ex": (C.c_int, [C.POINTER(WorkerConfig), C.POINTER(ResourceUsage)]),
    "mlsb_validate_device": (C.c_int, [C.c_int]),
    "mlsb_worker_stream": (_P, [_P]),
    "mlsb_worker_synchronize": (C.c_int, [_P]),
    "mlsb_host_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "mlsb_host_free": (C.c_int, [_P]),
    "mlsb_splats_upload_async": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32]),
    "mlsb_splats_device": (C.c_int, [_P, C.POINTER(_P)]),
    "mlsb_octree_build": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, _P, C.c_uint32]),
    "mlsb_octree_get": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_uint32), C.POINTER(_P), C.POINTER(C.c_uint32),
                                  C.POINTER(C.c_uint32)]),
    "mlsb_mls_set": (C.c_int, [_P, _P, C.c_uint32]),
    "mlsb_mls_set_buffers": (C.c_int, [_P, _P, _P, _P, _P, C.c_uint32]),
    "mlsb_mls_set_boundary_limit": (C.c_int, [_P, C.c_float]),
    "mlsb_mls_corners": (C.c_int, [_P, _P, C.c_uint32, C.POINTER(Swathe)]),
    "mlsb_set_scale_bias": (C.c_int, [_P, C.c_float, _P]),
    "mlsb_marching_generate": (C.c_int, [_P, _P, _P, MESH_SINK, _P]),
    "mlsb_marching_generate_host_field": (C.c_int, [_P, _P, _P, _P, MESH_SINK, _P]),
    "mlsb_marching_generate_cb": (C.c_int, [_P, FIELD_GENERATOR, _P, _P, _P, MESH_SINK, _P]),
    "mlsb_marching_generate_device": (C.c_int, [_P, FIELD_GENERATOR, _P, _P, _P, DEVICE_MESH_SINK, _P]),
    "mlsb_device_mesh_read": (C.c_int, [_P, C.POINTER(DeviceMesh), _P]),
    "mlsb_worker_add_mesh_filter": (C.c_int, [_P, MESH_FILTER, _P]),
    "mlsb_worker_clear_mesh_filters": (C.c_int, [_P]),
    "mlsb_mesh_filter_chain_apply": (C.c_int, [_P, C.POINTER(DeviceMesh), DEVICE_MESH_SINK, _P]),
    "mlsb_mesh_host_bytes": (C.c_uint64, [C.POINTER(MeshSizes)]),
    "mlsb_marching_tables": (C.c_int, [_P, _P, _P, _P]),
    "mlsb_worker_process_bin": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, _P, MESH_SINK, _P]),
    "mlsb_group_create": (C.c_int, [C.POINTER(WorkerConfig), C.c_uint32, C.c_uint32, C.c_int, GROUP_SINK, _P, C.POINTER(_P)]),
    "mlsb_group_destroy": (None, [_P]),
    "mlsb_group_submit": (C.c_int, [_P, _P, C.c_uint32, _P, _P, C.c_uint64]),
    "mlsb_group_wait": (C.c_int, [_P]),
    "mlsb_group_set_concurrent_sink": (C.c_int, [_P, C.c_int]),
    "mlsb_group_released": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "mlsb_group_worker": (C.c_int, [_P, C.c_uint32, C.POINTER(_P)]),
    "mlsb_timeplot_now": (C.c_double, []),
    "mlsb_group_set_timeplot": (C.c_int, [_P, C.c_int]),
    "mlsb_group_timeplot": (C.c_int, [_P, C.CFUNCTYPE(None, _P, C.c_char_p, C.c_char_p, C.c_double, C.c_double), _P]),
    "mlsb_group_counters": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "mlsb_worker_set_profiling": (C.c_int, [_P, C.c_int]),
    "mlsb_worker_get_statistic": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "mlsb_worker_visit_statistics": (C.c_int, [_P, C.CFUNCTYPE(None, _P, C.c_char_p, C.c_double, C.c_double, C.c_uint64), _P]),
    "mlsb_worker_reset_statistics": (C.c_int, [_P]),
    "mlsb_worker_set_work_counting": (C.c_int, [_P, C.c_int]),
    "mlsb_measure_fp32_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "mlsb_worker_get_mls_work": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "mlsb_worker_get_marching_work": (C.c_int, [_P] + [C.POINTER(C.c_uint64)] * 5),
    "mlsb_worker_launch_count": (C.c_uint64, [_P]),
    "mlsb_device_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "mlsb_device_free": (C.c_int, [_P]),
    "mlsb_memcpy_h2d": (C.c_int, [_P, _P, C.c_size_t]),
    "mlsb_memcpy_d2h": (C.c_int, [_P, _P, C.c_size_t]),
    "mlsb_device_memset": (C.c_int, [_P, C.c_int, C.c_size_t]),
    "mlsb_memcpy_d2d_async": (C.c_int, [_P, _P, C.c_size_t, _P]),
}

_lib = None


def library_path():
    # MLSB200_LIB selects another build of the same library (A/B runs of kernel variants)
    return os.environ.get("MLSB200_LIB") or os.path.join(_HERE, "libmlsb200.so")


def lib():
    """Loads libmlsb200.so (fails loudly if it has not been built: there is no fallback)."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise MlsbError(4, "%s is missing: run `make -C mlsgpu_b200` (or __graft_entry__.build())" % path)
        handle = C.CDLL(path)
        for name, (res, args) in API.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def _check(status):
    if status != 0:
        raise MlsbError(status, lib().mlsb_last_error_string().decode())


def version():
    return lib().mlsb_version().decode()


def device_count():
    n = C.c_int(0)
    _check(lib().mlsb_device_count(C.byref(n)))
    return n.value


def validate_device(device=0):
    """Raises MlsbError with the reason if the library cannot run on `device`."""
    _check(lib().mlsb_validate_device(device))


def resource_usage(**kw):
    """mlsb_worker_resource_usage_ex for a Worker(**kw) configuration: dict of device / largest allocation / pinned bytes."""
    cfg = Worker.config(**kw)
    u = ResourceUsage()
    _check(lib().mlsb_worker_resource_usage_ex(C.byref(cfg), C.byref(u)))
    return dict(device_bytes=u.deviceBytes, max_allocation_bytes=u.maxAllocationBytes, pinned_host_bytes=u.pinnedHostBytes)


def measure_fp32_peak(device=0):
    """FP32 lane-instructions per second of a dependent-free FFMA stream."""
    v = C.c_double()
    _check(lib().mlsb_measure_fp32_peak(device, C.byref(v)))
    return v.value


def marching_tables():
    count = np.zeros((256, 2), np.uint8)
    start = np.zeros((257, 2), np.uint16)
    data = np.zeros(8192, np.uint8)
    key = np.zeros((2432, 3), np.uint32)
    _check(lib().mlsb_marching_tables(count.ctypes.data, start.ctypes.data, data.ctypes.data, key.ctypes.data))
    return dict(count=count, start=start, data=data, key=key)


def _arr3(v, dtype):
    a = np.ascontiguousarray(v, dtype)
    assert a.shape == (3,)
    return a


class DeviceBuffer:
    """Raw device allocation (tests and benchmarks only)."""

    def __init__(self, nbytes):
        self.ptr = _P()
        self.nbytes = nbytes
        _check(lib().mlsb_device_alloc(C.byref(self.ptr), nbytes))

    @classmethod
    def from_array(cls, a):
        a = np.ascontiguousarray(a)
        buf = cls(a.nbytes)
        if a.nbytes:
            _check(lib().mlsb_memcpy_h2d(buf.ptr, a.ctypes.data, a.nbytes))
        return buf

    def fill(self, byte):
        _check(lib().mlsb_device_memset(self.ptr, byte, self.nbytes))

    def to_array(self, dtype, shape):
        out = np.empty(shape, dtype)
        assert out.nbytes <= self.nbytes
        if out.nbytes:
            _check(lib().mlsb_memcpy_d2h(out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            lib().mlsb_device_free(self.ptr)
            self.ptr = _P()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedBuffer:
    """Pinned host staging memory exposed as a numpy array (CLH::PinnedMemory)."""

    def __init__(self, dtype, count):
        self.ptr = _P()
        dtype = np.dtype(dtype)
        self.nbytes = dtype.itemsize * count
        _check(lib().mlsb_host_alloc(C.byref(self.ptr), self.nbytes))
        raw = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr.value)
        self.array = np.frombuffer(raw, dtype=dtype, count=count)

    def free(self):
        if self.ptr:
            self.array = None
            lib().mlsb_host_free(self.ptr)
            self.ptr = _P()


class Worker:
    """One worker = one GPU = one host thread (DeviceWorkerGroupBase::Worker)."""

    @classmethod
    def borrowed(cls, handle, cfg):
        """A view of a worker owned by a WorkerGroup (never destroyed from here)."""
        self = cls.__new__(cls)
        self.cfg = cfg
        self.h = _P(handle)
        self._keep = []
        self._borrowed = True
        return self

    @staticmethod
    def config(device=0, max_levels=6, subsampling=3, max_splats=1 << 20, max_cells=255,
               mesh_bytes=None, shape=0, boundary_limit=1.0):
        if mesh_bytes is None:
            mesh_bytes = max(64 << 20, max_cells * max_cells * 872)
        return WorkerConfig(device, max_levels, subsampling, max_splats, max_cells, mesh_bytes, shape, boundary_limit)

    def __init__(self, **kw):
        self.cfg = Worker.config(**kw)
        self.h = _P()
        _check(lib().mlsb_worker_create(C.byref(self.cfg), C.byref(self.h)))
        self._keep = []

    # ---- lifetime --------------------------------------------------------------------------
    def close(self):
        if self.h and not getattr(self, "_borrowed", False):
            lib().mlsb_worker_destroy(self.h)
        self.h = _P()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _check(lib().mlsb_worker_synchronize(self.h))

    @property
    def stream(self):
        return lib().mlsb_worker_stream(self.h)

    # ---- splats ----------------------------------------------------------------------------
    def upload_splats(self, splats, dst_first=0):
        """Async H2D from a host array (pinned or pageable); keep the array alive until the
        next synchronize()."""
        assert splats.dtype == SPLAT_DTYPE and splats.flags.c_contiguous
        self._keep.append(splats)
        _check(lib().mlsb_splats_upload_async(self.h, splats.ctypes.data, dst_first, len(splats)))

    def upload_splats_from_device(self, src, src_first, n, dst_first=0):
        """Device-to-device refill of the splat buffer (benchmarks with HBM-resident inputs)."""
        _check(lib().mlsb_splats_upload_async(self.h, src.ptr.value + 32 * src_first, dst_first, n))

    def download_splats(self, first, n):
        self.synchronize()
        ptr = _P()
        _check(lib().mlsb_splats_device(self.h, C.byref(ptr)))
        out = np.empty(n, SPLAT_DTYPE)
        if n:
            _check(lib().mlsb_memcpy_d2h(out.ctypes.data, ptr.value + 32 * first, out.nbytes))
        return out

    # ---- octree ----------------------------------------------------------------------------
    def octree_build(self, first, n, size, offset, subsampling=None):
        size = _arr3(size, np.uint32)
        offset = _arr3(offset, np.int32)
        sub = self.cfg.subsampling if subsampling is None else subsampling
        _check(lib().mlsb_octree_build(self.h, first, n, size.ctypes.data, offset.ctypes.data, sub))

    def octree_get(self, num_commands=None):
        """Host copies of (start, commands)."""
        self.synchronize()
        start, commands = _P(), _P()
        ns, nc, nl = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().mlsb_octree_get(self.h, C.byref(start), C.byref(ns), C.byref(commands), C.byref(nc), C.byref(nl)))
        hs = np.empty(ns.value, np.int32)
        _check(lib().mlsb_memcpy_d2h(hs.ctypes.data, start, hs.nbytes))
        n = nc.value if num_commands is None else min(num_commands, nc.value)
        hc = np.empty(n, np.int32)
        if n:
            _check(lib().mlsb_memcpy_d2h(hc.ctypes.data, commands, hc.nbytes))
        return hs, hc

    def octree_fill_commands(self, byte):
        """Overwrites the command buffer (tests: detect slots the build does not write)."""
        start, commands = _P(), _P()
        ns, nc, nl = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().mlsb_octree_get(self.h, C.byref(start), C.byref(ns), C.byref(commands), C.byref(nc), C.byref(nl)))
        _check(lib().mlsb_device_memset(commands, byte, 4 * nc.value))

    # ---- MLS -------------------------------------------------------------------------------
    def mls_set(self, offset, subsampling=None):
        offset = _arr3(offset, np.int32)
        sub = self.cfg.subsampling if subsampling is None else subsampling
        _check(lib().mlsb_mls_set(self.h, offset.ctypes.data, sub))

    def mls_set_buffers(self, offset, splats, commands, start, subsampling):
        offset = _arr3(offset, np.int32)
        _check(lib().mlsb_mls_set_buffers(self.h, offset.ctypes.data, splats.ptr, commands.ptr, start.ptr, subsampling))

    def mls_set_boundary_limit(self, limit):
        _check(lib().mlsb_mls_set_boundary_limit(self.h, limit))

    def mls_corners(self, field, pitch, width, height, z_stride, z_bias, z_first, z_last):
        sw = Swathe(width, height, z_stride, z_bias, z_first, z_last)
        _check(lib().mlsb_mls_corners(self.h, field.ptr, pitch, C.byref(sw)))

    # ---- marching --------------------------------------------------------------------------
    def set_scale_bias(self, scale, bias):
        bias = _arr3(bias, np.float32)
        _check(lib().mlsb_set_scale_bias(self.h, scale, bias.ctypes.data))

    @staticmethod
    def _collector(blocks):
        def sink(sizes_p, blob, _user):
            s = sizes_p.contents
            nv, nt, ni = int(s.numVertices), int(s.numTriangles), int(s.numInternalVertices)
            ne = nv - ni
            nbytes = 8 * ne + 12 * nv + 12 * nt
            raw = np.frombuffer((C.c_char * max(nbytes, 1)).from_address(blob), np.uint8, nbytes).copy()
            keys = raw[:8 * ne].view(np.uint64)
            verts = raw[8 * ne:8 * ne + 12 * nv].view(np.float32).reshape(nv, 3)
            tris = raw[8 * ne + 12 * nv:].view(np.uint32).reshape(nt, 3)
            blocks.append(dict(vertices=verts, external_keys=keys, triangles=tris, num_internal=ni))
        return MESH_SINK(sink)

    def marching_generate(self, size, key_offset):
        """Marching::generate with the bound MlsFunctor; returns the shipped-out blocks."""
        blocks = []
        cb = self._collector(blocks)
        size = _arr3(size, np.uint32)
        ko = _arr3(key_offset, np.uint32)
        _check(lib().mlsb_marching_generate(self.h, size.ctypes.data, ko.ctypes.data, cb, None))
        return blocks

    def marching_generate_host_field(self, field, key_offset):
        """field: [depth][height][width] float32 volume."""
        field = np.ascontiguousarray(field, np.float32)
        d, h, w = field.shape
        blocks = []
        cb = self._collector(blocks)
        size = _arr3((w, h, d), np.uint32)
        ko = _arr3(key_offset, np.uint32)
        _check(lib().mlsb_marching_generate_host_field(self.h, field.ctypes.data, size.ctypes.data, ko.ctypes.data,
                                                       cb, None))
        return blocks

    def marching_generate_cb(self, fill, size, key_offset):
        """Marching::generate with a generator callback: fill(device_field_ptr, pitch, swathe, stream) -> 0."""
        blocks = []
        cb = self._collector(blocks)
        gen = FIELD_GENERATOR(lambda _u, field, pitch, sw, stream: int(fill(field, pitch, sw.contents, stream) or 0))
        size = _arr3(size, np.uint32)
        ko = _arr3(key_offset, np.uint32)
        _check(lib().mlsb_marching_generate_cb(self.h, gen, None, size.ctypes.data, ko.ctypes.data, cb, None))
        return blocks

    def marching_generate_device(self, size, key_offset, visit=None, fill=None):
        """Marching::generate handing out DeviceKeyMesh blocks.  visit(mesh: DeviceMesh, stream) may change the
        device buffers (a MeshFilter); every block is then read back with mlsb_device_mesh_read."""
        blocks = []

        def sink(mesh_p, stream, _user):
            mesh = mesh_p.contents
            if visit is not None:
                visit(mesh, stream)
            s = mesh.sizes
            nv, nt, ni = int(s.numVertices), int(s.numTriangles), int(s.numInternalVertices)
            ne = nv - ni
            raw = np.empty(max(8 * ne + 12 * nv + 12 * nt, 1), np.uint8)
            st = lib().mlsb_device_mesh_read(self.h, mesh_p, raw.ctypes.data)
            if st != 0:
                return st
            blocks.append(dict(vertices=raw[8 * ne:8 * ne + 12 * nv].view(np.float32).reshape(nv, 3).copy(),
                               external_keys=raw[:8 * ne].view(np.uint64).copy(),
                               triangles=raw[8 * ne + 12 * nv:8 * ne + 12 * nv + 12 * nt].view(np.uint32).reshape(nt, 3).copy(),
                               num_internal=ni))
            return 0
        cb = DEVICE_MESH_SINK(sink)
        gen = FIELD_GENERATOR(lambda _u, field, pitch, sw, stream: int(fill(field, pitch, sw.contents, stream) or 0)) \
            if fill is not None else FIELD_GENERATOR()
        size = _arr3(size, np.uint32)
        ko = _arr3(key_offset, np.uint32)
        _check(lib().mlsb_marching_generate_device(self.h, gen, None, size.ctypes.data, ko.ctypes.data, cb, None))
        return blocks

    def add_mesh_filter(self, fn):
        """MeshFilterChain::addFilter: fn(in_mesh: DeviceMesh, out_mesh: DeviceMesh, stream) -> 0 / None on success.
        out_mesh starts as a copy of in_mesh; device work goes on `stream`."""
        cb = MESH_FILTER(lambda _u, i, o, stream: int(fn(i.contents, o.contents, stream) or 0))
        self._filters = getattr(self, "_filters", []) + [cb]      # keep the thunks alive
        _check(lib().mlsb_worker_add_mesh_filter(self.h, cb, None))

    def clear_mesh_filters(self):
        _check(lib().mlsb_worker_clear_mesh_filters(self.h))
        self._filters = []

    def read_device_mesh(self, mesh):
        """enqueueReadMesh into numpy arrays: dict(vertices, external_keys, triangles, num_internal)."""
        s = mesh.sizes
        nv, nt, ni = int(s.numVertices), int(s.numTriangles), int(s.numInternalVertices)
        ne = nv - ni
        raw = np.empty(max(8 * ne + 12 * nv + 12 * nt, 1), np.uint8)
        _check(lib().mlsb_device_mesh_read(self.h, C.byref(mesh), raw.ctypes.data))
        return dict(vertices=raw[8 * ne:8 * ne + 12 * nv].view(np.float32).reshape(nv, 3).copy(),
                    external_keys=raw[:8 * ne].view(np.uint64).copy(),
                    triangles=raw[8 * ne + 12 * nv:8 * ne + 12 * nv + 12 * nt].view(np.uint32).reshape(nt, 3).copy(),
                    num_internal=ni)

    def mesh_filter_chain_apply(self, mesh):
        """MeshFilterChain::operator() on a caller's DeviceMesh: the filtered mesh, read back."""
        out = []

        def sink(mesh_p, _stream, _user):
            out.append(self.read_device_mesh(mesh_p.contents))
            return 0
        cb = DEVICE_MESH_SINK(sink)
        _check(lib().mlsb_mesh_filter_chain_apply(self.h, C.byref(mesh), cb, None))
        return out[0]

    def process_bin(self, first, n, size, offset, collect=True):
        """DeviceWorkerGroupBase::Worker::operator() for one bin."""
        blocks = []
        cb = self._collector(blocks) if collect else MESH_SINK(lambda a, b, c: None)
        size = _arr3(size, np.uint32)
        offset = _arr3(offset, np.int32)
        _check(lib().mlsb_worker_process_bin(self.h, first, n, size.ctypes.data, offset.ctypes.data, cb, None))
        self._keep.clear()
        return blocks

    # ---- statistics ------------------------------------------------------------------------
    def set_profiling(self, level):
        """0 off, 1 processCorners only, 2 (or True) every kernel."""
        _check(lib().mlsb_worker_set_profiling(self.h, 2 if level is True else int(level)))

    def statistic(self, name):
        s, c = C.c_double(), C.c_uint64()
        _check(lib().mlsb_worker_get_statistic(self.h, name.encode(), C.byref(s), C.byref(c)))
        return s.value, c.value

    def statistics(self):
        """{name: (sum, sum of squares, count)} of every statistic with samples (kernel times need set_profiling)."""
        out = {}
        visitor = C.CFUNCTYPE(None, _P, C.c_char_p, C.c_double, C.c_double, C.c_uint64)(
            lambda _u, name, s, s2, c: out.__setitem__(name.decode(), (s, s2, c)))
        _check(lib().mlsb_worker_visit_statistics(self.h, visitor, None))
        return out

    def reset_statistics(self):
        _check(lib().mlsb_worker_reset_statistics(self.h))

    def set_work_counting(self, enabled):
        _check(lib().mlsb_worker_set_work_counting(self.h, int(enabled)))

    def mls_work(self):
        t, h = C.c_uint64(), C.c_uint64()
        _check(lib().mlsb_worker_get_mls_work(self.h, C.byref(t), C.byref(h)))
        return t.value, h.value

    def marching_work(self):
        """(corners, occupied cells, unwelded vertices, indices, welded vertices) since the last reset."""
        v = [C.c_uint64() for _ in range(5)]
        _check(lib().mlsb_worker_get_marching_work(self.h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    @property
    def launch_count(self):
        return int(lib().mlsb_worker_launch_count(self.h))


class WorkerGroup:
    """CopyGroup + DeviceWorkerGroup: `num_workers` native worker threads on one GPU fed by a bounded
    queue of bins (mlsb_group_*).  `sink(tag, block)` is called, serialised, with each mesh block
    (same dict layout as Worker.marching_generate); sink=None only counts bytes."""

    def __init__(self, num_workers=4, queue_depth=0, stage=True, sink=None, device=0, max_levels=6, subsampling=3,
                 max_splats=1 << 20, max_cells=255, mesh_bytes=None, shape=0, boundary_limit=1.0):
        if mesh_bytes is None:
            mesh_bytes = max(64 << 20, max_cells * max_cells * 872)
        self.cfg = WorkerConfig(device, max_levels, subsampling, max_splats, max_cells, mesh_bytes, shape, boundary_limit)
        self.num_workers = num_workers
        self._keep = []
        if sink is not None:
            def cb(tag, sizes_p, blob, _user):
                blocks = []
                Worker._collector(blocks)(sizes_p, blob, None)
                sink(int(tag), blocks[0])
            self._cb = GROUP_SINK(cb)
        else:
            self._cb = GROUP_SINK()
        self.h = _P()
        _check(lib().mlsb_group_create(C.byref(self.cfg), num_workers, queue_depth, int(stage), self._cb, None,
                                       C.byref(self.h)))

    def worker(self, i):
        h = _P()
        _check(lib().mlsb_group_worker(self.h, i, C.byref(h)))
        return Worker.borrowed(h.value, self.cfg)

    @property
    def workers(self):
        return [self.worker(i) for i in range(self.num_workers)]

    def submit(self, splats, size, offset, tag=0):
        """splats: numpy array (host) — kept alive until wait() — or (device_ptr, count)."""
        size = _arr3(size, np.uint32)
        offset = _arr3(offset, np.int32)
        if isinstance(splats, tuple):
            ptr, n = splats
        else:
            assert splats.dtype == SPLAT_DTYPE and splats.flags.c_contiguous
            self._keep.append(splats)
            ptr, n = splats.ctypes.data, len(splats)
        _check(lib().mlsb_group_submit(self.h, ptr, n, size.ctypes.data, offset.ctypes.data, tag))

    def wait(self):
        _check(lib().mlsb_group_wait(self.h))
        self._keep.clear()

    def released(self):
        """Leading submits whose splat memory the group will not read again (mlsb_group_released)."""
        n = C.c_uint64()
        _check(lib().mlsb_group_released(self.h, C.byref(n)))
        return n.value

    def set_concurrent_sink(self, enabled=True):
        _check(lib().mlsb_group_set_concurrent_sink(self.h, int(enabled)))

    def set_timeplot(self, enabled=True):
        _check(lib().mlsb_group_set_timeplot(self.h, int(enabled)))

    def timeplot(self):
        """[(worker, action, t0, t1)] recorded since set_timeplot (Timeplot::Action events)."""
        out = []
        cb = C.CFUNCTYPE(None, _P, C.c_char_p, C.c_char_p, C.c_double, C.c_double)(
            lambda _u, w, a, t0, t1: out.append((w.decode(), a.decode(), t0, t1)))
        _check(lib().mlsb_group_timeplot(self.h, cb, None))
        return out

    def counters(self):
        b, s, m = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().mlsb_group_counters(self.h, C.byref(b), C.byref(s), C.byref(m)))
        return b.value, s.value, m.value

    def close(self):
        if self.h:
            lib().mlsb_group_destroy(self.h)
            self.h = _P()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
