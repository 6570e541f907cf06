"""CPU-side checks of the boundary: the library loads, exports every symbol that
include/mlsb200.h declares, and refuses to compute without a GPU (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import mlsgpu_b200
from mlsgpu_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mlsb200.h")).read()
    return sorted(set(re.findall(r"MLSB_API\s+[^;(]*?\b(mlsb_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    syms = header_symbols()
    assert len(syms) >= 30
    assert sorted(capi.API) == syms


def test_library_exports_every_symbol():
    path = mlsgpu_b200.library_path()
    assert os.path.exists(path), "build the library first: make -C mlsgpu_b200"
    handle = ctypes.CDLL(path)
    for name in header_symbols():
        assert hasattr(handle, name), name
    assert "sm_100a" in mlsgpu_b200.version()


def test_marching_tables_match_oracle():
    from oracle.oracle import Oracle
    mine = mlsgpu_b200.marching_tables()
    ref = Oracle().marching_tables()
    for k in ("count", "start", "data", "key"):
        np.testing.assert_array_equal(mine[k], ref[k])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mlsgpu_b200.MlsbError):
        mlsgpu_b200.Worker(max_splats=1024, max_cells=15)


def test_host_binning_covers_every_splat():
    from mlsgpu_b200 import host
    s = host.sphere_splats(20000, seed=3, radius=0.02)
    g, bins, lo, spacing = host.prepare(s, 0.02, max_cells=63, max_splats=3000)
    assert len(bins) >= 8
    seen = np.zeros(len(g), bool)
    for b in bins:
        assert (b["size"] <= 64).all() and len(b["ids"]) <= 3000
        seen[b["ids"]] = True
        p, r = g["position"][b["ids"]], g["radius"][b["ids"]]
        o, sz = b["offset"].astype(np.float32), b["size"].astype(np.float32)
        assert ((p + r[:, None] >= o) & (p - r[:, None] <= o + sz - 1)).all()
    assert seen.all()
    assert np.allclose(g["quality"], 1.0 / (0.02 * 4.0) ** 2, rtol=1e-6)
