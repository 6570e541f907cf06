"""The parts of the boundary beyond the per-bin call: Marching::Generator as an open interface
(src/marching.h:204-253), DeviceKeyMesh / enqueueReadMesh and a device-side MeshFilter hook
(src/mesh.h:101-180, src/mesh_filter.h:55-116), validateDevice and resourceUsage
(src/splat_tree_cl.h:216-232, src/marching.h:143-171, src/workers.cpp:176-205)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.oracle import Oracle
from gpu_common import assert_block_equal
from test_oracle_marching import sphere_field, to_image


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def test_validate_device_and_limits():
    import mlsgpu_b200 as m
    from mlsgpu_b200 import capi
    capi.validate_device(0)
    with pytest.raises(m.MlsbError):
        capi.validate_device(m.device_count() + 3)
    assert (capi.MAX_LEVELS, capi.MAX_SPLATS) == (10, (1 << 31) // 16)          # src/splat_tree_cl.h:76,88
    with pytest.raises(m.MlsbError):
        m.Worker(max_levels=11, max_splats=16, max_cells=31)
    with pytest.raises(m.MlsbError):
        m.Worker(max_splats=capi.MAX_SPLATS + 1, max_cells=31)
    with pytest.raises(m.MlsbError):
        m.Worker(max_splats=16, max_cells=600)                                  # above MAX_DIMENSION as built
    with pytest.raises(m.MlsbError):
        m.Worker(max_splats=16, max_cells=255, mesh_bytes=1 << 20)              # less than one slice of cells


@pytest.mark.parametrize("kw", [dict(max_splats=1 << 18, max_cells=127, mesh_bytes=32 << 20),
                                dict(max_splats=1 << 21, max_cells=255, mesh_bytes=256 << 20),
                                dict(max_splats=1 << 16, max_cells=63, max_levels=8, mesh_bytes=16 << 20)])
def test_resource_usage_matches_what_a_worker_allocates(kw):
    """resourceUsage is what a caller sizes --device-threads by: it must not under-estimate, and not
    over-estimate by more than the allocator's granularity."""
    import torch
    import mlsgpu_b200 as m
    from mlsgpu_b200 import capi
    torch.cuda.init()
    with m.Worker(max_splats=16, max_cells=31):     # context, module and first-allocation overheads are paid here
        pass
    torch.cuda.synchronize()
    want = capi.resource_usage(**kw)
    free0, _ = torch.cuda.mem_get_info(0)
    with m.Worker(**kw) as w:
        free1, _ = torch.cuda.mem_get_info(0)
        used = free0 - free1
        assert w.launch_count == 0
    slack = 40 * (2 << 20)                          # ~30 allocations, each rounded up to the 2 MiB granule
    assert want["device_bytes"] <= used + (4 << 20), (want, used)
    assert used <= want["device_bytes"] + slack, (want, used)
    assert want["max_allocation_bytes"] <= want["device_bytes"] and want["pinned_host_bytes"] > 0


def upload_volume(vol):
    """A Marching::Generator (the HostGenerator of test/test_marching.cpp:62-129): copies a host volume
    into the worker's field, row by row at the pitch it is given."""
    from mlsgpu_b200.capi import lib
    calls = []

    def fill(field, pitch, sw, stream):
        d, h, w_ = vol.shape
        assert (sw.width, sw.height, sw.zFirst, sw.zLast) == (w_, h, 0, d - 1) and pitch >= w_
        calls.append((pitch, sw.zStride))
        row = np.full(pitch, np.nan, np.float32)
        for z in range(d):
            for y in range(h):
                row[:w_] = vol[z, y]
                assert lib().mlsb_memcpy_h2d(field + 4 * pitch * (y + z * sw.zStride + sw.zBias), row.ctypes.data, 4 * pitch) == 0
        return 0
    return fill, calls


def test_generator_callback_equals_host_field(orc):
    import mlsgpu_b200 as m
    vol = sphere_field(40, 36, 33, 19.5, 17.0, 16.0, 12.0)
    vol[10, 5, 7] = np.nan
    img, zs = to_image(vol)
    ref = orc.marching_generate(img, 40, 36, zs, 0, 0, 32, (3, 4, 5))
    fill, calls = upload_volume(vol)
    with m.Worker(max_splats=16, max_cells=63) as w:
        blocks = w.marching_generate_cb(fill, (40, 36, 33), (3, 4, 5))
        direct = w.marching_generate_host_field(vol, (3, 4, 5))
        with pytest.raises(m.MlsbError):
            w.marching_generate_cb(lambda *a: 7, (40, 36, 33), (3, 4, 5))       # a failing generator aborts the generate
    assert len(calls) == 1 and len(blocks) == 1
    assert_block_equal(blocks[0], ref)
    assert_block_equal(direct[0], ref)


def test_device_mesh_sink_and_filter(orc):
    """DeviceKeyMesh blocks: read back with mlsb_device_mesh_read they equal the host path; a sink that
    rewrites a vertex on the device first (a MeshFilter) sees its change in what is read back."""
    import mlsgpu_b200 as m
    from mlsgpu_b200.capi import lib
    vol = sphere_field(40, 36, 33, 19.5, 17.0, 16.0, 21.0)       # cut by the region: internal and external vertices
    img, zs = to_image(vol)
    ref = orc.marching_generate(img, 40, 36, zs, 0, 0, 32, (0, 0, 0))
    assert 0 < ref["num_internal"] < len(ref["vertices"])
    fill, _ = upload_volume(vol)
    marker = np.array([1.5, -2.5, 3.5], np.float32)
    seen = []

    def move_last_vertex(mesh, stream):
        seen.append((int(mesh.sizes.numVertices), int(mesh.sizes.numInternalVertices)))
        last = mesh.vertices + 12 * (int(mesh.sizes.numVertices) - 1)
        assert lib().mlsb_memcpy_h2d(last, marker.ctypes.data, 12) == 0
    with m.Worker(max_splats=16, max_cells=63) as w:
        plain = w.marching_generate_device((40, 36, 33), (0, 0, 0), fill=fill)
        filtered = w.marching_generate_device((40, 36, 33), (0, 0, 0), visit=move_last_vertex, fill=fill)
    assert len(plain) == len(filtered) == 1
    assert_block_equal(plain[0], ref)
    assert seen == [(len(ref["vertices"]), ref["num_internal"])]
    np.testing.assert_array_equal(filtered[0]["vertices"][-1], marker)
    np.testing.assert_array_equal(filtered[0]["vertices"][:-1], ref["vertices"][:-1])
    np.testing.assert_array_equal(filtered[0]["triangles"], ref["triangles"])
    np.testing.assert_array_equal(filtered[0]["external_keys"], ref["keys"][ref["num_internal"]:])


def test_device_mesh_from_bound_mls(orc):
    """generate == NULL: the MlsFunctor bound by mlsb_mls_set is the generator, as in process_bin."""
    import mlsgpu_b200 as m
    from mlsgpu_b200 import host
    s = host.sphere_splats(20000, seed=1, radius=0.02, noise=0.004)
    g, bins, lo, spacing = host.prepare(s, 0.04)
    b = bins[0]
    splats = np.ascontiguousarray(g[b["ids"]])
    with m.Worker(max_splats=len(splats), max_cells=int(b["size"].max()) - 1) as w:
        w.upload_splats(splats)
        want = w.process_bin(0, len(splats), b["size"], b["offset"])
        w.upload_splats(splats)
        w.octree_build(0, len(splats), [-(-int(v) // 8) * 8 for v in b["size"]], b["offset"])
        w.mls_set(b["offset"])
        got = w.marching_generate_device(b["size"], b["offset"].astype(np.uint32))
    assert len(want) == len(got) == 1
    for k in ("vertices", "triangles", "external_keys"):
        np.testing.assert_array_equal(got[0][k], want[0][k])


class TruncateFilter:
    """The test filter of test/test_mesh_filter.cpp:42-108: copies the mesh into its own device buffers,
    dropping the FIRST triangle."""

    def __init__(self, max_vertices, max_triangles):
        import torch
        self.vertices = torch.empty(3 * max_vertices, dtype=torch.float32, device="cuda")
        self.keys = torch.empty(max_vertices, dtype=torch.int64, device="cuda")
        self.triangles = torch.empty(3 * max_triangles, dtype=torch.int32, device="cuda")
        self.calls = 0

    def __call__(self, inp, out, stream):
        from mlsgpu_b200.capi import lib
        self.calls += 1
        nv, nt = int(inp.sizes.numVertices), int(inp.sizes.numTriangles)
        assert inp.vertices != self.vertices.data_ptr() and inp.triangles != self.triangles.data_ptr()
        assert 3 * nv <= self.vertices.numel() and 3 * nt <= self.triangles.numel()
        assert lib().mlsb_memcpy_d2d_async(self.vertices.data_ptr(), inp.vertices, 12 * nv, stream) == 0
        assert lib().mlsb_memcpy_d2d_async(self.keys.data_ptr(), inp.vertexKeys, 8 * nv, stream) == 0
        if nt > 0:
            assert lib().mlsb_memcpy_d2d_async(self.triangles.data_ptr(), inp.triangles + 12, 12 * (nt - 1), stream) == 0
        out.vertices, out.vertexKeys, out.triangles = self.vertices.data_ptr(), self.keys.data_ptr(), self.triangles.data_ptr()
        out.sizes.numTriangles = max(nt - 1, 0)
        return 0


def five_vertex_mesh():
    """test/test_mesh_filter.cpp:179-192: 5 vertices (4 internal), 3 triangles."""
    import torch
    from mlsgpu_b200 import capi
    v = torch.arange(15, dtype=torch.float32, device="cuda")
    k = torch.tensor([100, 200, 300, 400, 500], dtype=torch.int64, device="cuda")
    t = torch.tensor([0, 1, 2, 1, 2, 3, 2, 3, 4], dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    mesh = capi.DeviceMesh(capi.MeshSizes(5, 3, 4), v.data_ptr(), k.data_ptr(), t.data_ptr())
    return mesh, (v, k, t)


def test_mesh_filter_chain_reference_cases():
    """TestMeshFilterChain (test/test_mesh_filter.cpp:204-268): no filters, two filters, a chain that empties the
    mesh (still handed over), an empty mesh going in."""
    import mlsgpu_b200 as m
    with m.Worker(max_splats=16, max_cells=31) as w:
        mesh, keep = five_vertex_mesh()
        got = w.mesh_filter_chain_apply(mesh)                                   # testNoFilters
        assert (len(got["vertices"]), got["num_internal"], len(got["triangles"])) == (5, 4, 3)
        np.testing.assert_array_equal(got["vertices"].ravel(), np.arange(15, dtype=np.float32))
        np.testing.assert_array_equal(got["external_keys"], [500])
        f1, f2 = TruncateFilter(10, 10), TruncateFilter(10, 10)
        w.add_mesh_filter(f1)
        w.add_mesh_filter(f2)
        got = w.mesh_filter_chain_apply(mesh)                                   # testFilters
        assert (len(got["vertices"]), got["num_internal"], len(got["triangles"])) == (5, 4, 1)
        np.testing.assert_array_equal(got["triangles"], [[2, 3, 4]])            # the first two were dropped, in order
        np.testing.assert_array_equal(got["external_keys"], [500])
        w.add_mesh_filter(f1)
        got = w.mesh_filter_chain_apply(mesh)                                   # testCull
        assert (len(got["vertices"]), len(got["external_keys"]), len(got["triangles"])) == (5, 1, 0)
        empty, keep2 = five_vertex_mesh()
        empty.sizes = type(empty.sizes)(0, 0, 0)
        got = w.mesh_filter_chain_apply(empty)                                  # testEmpty
        assert (len(got["vertices"]), got["num_internal"], len(got["triangles"])) == (0, 0, 0)
        assert (f1.calls, f2.calls) == (1 + 2 + 2, 1 + 1 + 1)
        w.add_mesh_filter(lambda i, o, s: 3)                                    # a failing filter aborts
        with pytest.raises(m.MlsbError):
            w.mesh_filter_chain_apply(mesh)
        w.clear_mesh_filters()
        got = w.mesh_filter_chain_apply(mesh)
        assert len(got["triangles"]) == 3


def check_truncated(cut, plain, f):
    assert len(cut) == len(plain) == f.calls
    for c, p in zip(cut, plain):
        np.testing.assert_array_equal(c["vertices"], p["vertices"])
        np.testing.assert_array_equal(c["external_keys"], p["external_keys"])
        np.testing.assert_array_equal(c["triangles"], p["triangles"][1:])
        assert c["num_internal"] == p["num_internal"]


def test_mesh_filters_apply_to_every_block():
    """Filters installed on a worker sit between marching and the host sink (the place of DeviceWorkerGroup's
    filterChain, src/workers.cpp:224,263-274): each block arrives without its first triangle, everything else
    unchanged — for process_bin, and for a generate call that ships several blocks."""
    import mlsgpu_b200 as m
    from mlsgpu_b200 import host
    from test_oracle_marching import MARCHING_CASES
    s = host.sphere_splats(20000, seed=1, radius=0.02, noise=0.004)
    g, bins, lo, spacing = host.prepare(s, 0.04)
    b = bins[0]
    splats = np.ascontiguousarray(g[b["ids"]])
    with m.Worker(max_splats=len(splats), max_cells=int(b["size"].max()) - 1) as w:
        w.upload_splats(splats)
        plain = w.process_bin(0, len(splats), b["size"], b["offset"])
        f = TruncateFilter(len(plain[0]["vertices"]), len(plain[0]["triangles"]))
        w.add_mesh_filter(f)
        w.upload_splats(splats)
        cut = w.process_bin(0, len(splats), b["size"], b["offset"])
    assert len(plain) == 1
    check_truncated(cut, plain, f)
    vol = MARCHING_CASES["truncated"]()
    with m.Worker(max_splats=16, max_cells=82, mesh_bytes=82 * 82 * 872) as w:      # the minimum: several blocks
        plain = w.marching_generate_host_field(vol, (0, 0, 0))
        f = TruncateFilter(max(len(p["vertices"]) for p in plain), max(len(p["triangles"]) for p in plain))
        w.add_mesh_filter(f)
        cut = w.marching_generate_host_field(vol, (0, 0, 0))
    assert len(plain) > 1
    check_truncated(cut, plain, f)
