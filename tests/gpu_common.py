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
