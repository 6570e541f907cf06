"""Helpers for the -m gpu parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import numpy as np

from fixtures import walk_chain  # noqa: F401


def canonical_triangles(tris):
    """Rotation-normalised index triples, sorted (the reference's triangle order is arbitrary)."""
    t = np.asarray(tris, np.int64).reshape(-1, 3)
    if len(t) == 0:
        return t
    k = np.argmin(t, axis=1)
    rows = np.arange(len(t))
    out = np.stack([t[rows, k], t[rows, (k + 1) % 3], t[rows, (k + 2) % 3]], axis=1)
    order = np.lexsort((out[:, 2], out[:, 1], out[:, 0]))
    return out[order]


def assert_block_equal(gpu, orc, scale=1.0, bias=(0.0, 0.0, 0.0), exact=True):
    """gpu: block dict from Worker; orc: dict from Oracle.marching_generate (grid coordinates)."""
    nv = len(orc["vertices"])
    assert len(gpu["vertices"]) == nv
    assert gpu["num_internal"] == orc["num_internal"]
    want = orc["vertices"].copy()
    if scale != 1.0 or any(b != 0.0 for b in bias):
        b = np.asarray(bias, np.float32)
        want = (want.astype(np.float64) * np.float64(np.float32(scale)) + b.astype(np.float64))
        # fma(v, scale, bias) is a single rounding of the exact result
        want = want.astype(np.float32)
    if exact:
        np.testing.assert_array_equal(gpu["vertices"], want)
    else:
        np.testing.assert_allclose(gpu["vertices"], want, atol=1e-4)
    ni = orc["num_internal"]
    np.testing.assert_array_equal(gpu["external_keys"], orc["keys"][ni:])
    np.testing.assert_array_equal(gpu["triangles"], orc["triangles"])


def same_field(a, b):
    """NaN positions equal and finite values bit-identical."""
    na, nb = np.isnan(a), np.isnan(b)
    np.testing.assert_array_equal(na, nb)
    np.testing.assert_array_equal(a[~na].view(np.uint32), b[~nb].view(np.uint32))


def padded(size):
    """(pitch, zStride, depth): the bin's corners per axis rounded up to the 8-wide MLS blocks."""
    return tuple(-(-int(v) // 8) * 8 for v in size)


def gpu_field(w, splats, size, offset):
    """Octree build + processCorners of one bin on worker `w`, stage by stage through the C ABI.
    Returns (field image, start, commands, splats as mutated by the build, (T, H))."""
    from mlsgpu_b200.capi import DeviceBuffer
    pitch, zs, depth = padded(size)
    w.set_work_counting(True)
    w.reset_statistics()
    w.upload_splats(splats)
    w.octree_build(0, len(splats), padded(size), offset)
    w.mls_set(offset)
    d_img = DeviceBuffer(4 * pitch * zs * depth)
    d_img.fill(0)
    w.mls_corners(d_img, pitch, int(size[0]), int(size[1]), zs, 0, 0, int(size[2]) - 1)
    w.synchronize()
    got = d_img.to_array(np.float32, (zs * depth, pitch))
    d_img.free()
    start, commands = w.octree_get()
    out = w.download_splats(0, len(splats))
    work = w.mls_work()
    w.set_work_counting(False)
    return got, start, commands, out, work


def oracle_field(orc, splats, size, offset, limit=1.0, shape=0, max_levels=6, subsampling=3):
    """The same two stages on the oracle.  Returns (field image, tree, mutated splats, (T, H))."""
    s = splats.copy()
    tree = orc.octree_build(s, 0, len(s), offset, subsampling, max_levels)
    pitch, zs, depth = padded(size)
    img = np.zeros((zs * depth, pitch), np.float32)
    work = orc.mls_process_corners(img, s, tree["commands"], tree["start"], 3 * subsampling, offset, int(size[0]),
                                   int(size[1]), zs, 0, 0, int(size[2]) - 1, orc.boundary_factor(limit), shape)
    return img, tree, s, work


def oracle_blocks(orc, img, size, offset, vertex_space, index_space, z0=0, z1=None):
    """Marching on the oracle with the worker's overflow rule: a z range whose per-cell vertex or index
    count exceeds the mesh memory is halved and shipped as two blocks (shipRange in csrc/marching.cu;
    the reference splits by its per-slice histogram, src/marching.cpp:651-704)."""
    zs = padded(size)[1]
    z1 = int(size[2]) - 1 if z1 is None else z1
    mesh = orc.marching_generate(img, int(size[0]), int(size[1]), zs, 0, z0, z1, [int(v) for v in offset],
                                 vcap=8 << 20, icap=48 << 20)
    _, nv, ni = mesh["stats"]
    if nv > vertex_space or ni > index_space:
        assert z1 - z0 > 1
        mid = z0 + (z1 - z0) // 2
        return (oracle_blocks(orc, img, size, offset, vertex_space, index_space, z0, mid)
                + oracle_blocks(orc, img, size, offset, vertex_space, index_space, mid, z1))
    return [mesh] if len(mesh["vertices"]) else []


def mesh_space(mesh_bytes):
    """(vertexSpace, indexSpace) of a worker: meshBytes / 872 cells x 13 / x 36 (src/marching.cpp:389-392)."""
    cells = mesh_bytes // 872
    return cells * 13, cells * 36
