"""The C++ mirror of the reference classes (mlsgpu_b200/host/mlsb200.hpp): compiles on its own,
and on a GPU drives one bin to the same mesh as the Python binding."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "mlsgpu_b200", "host")


def test_mirror_header_compiles(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text('#include "mlsb200.hpp"\nint main() { return sizeof(mlsb::Splat) == 32 ? 0 : 1; }\n')
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-I", HOST, str(src)], check=True)


def fnv1a(data):
    h = 1469598103934665603
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.gpu
def test_mirror_demo_matches_python_binding(tmp_path):
    import mlsgpu_b200 as m
    from mlsgpu_b200 import host
    demo = os.path.join(HOST, "mirror_demo")
    if not os.path.exists(demo):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "mlsgpu_b200"), "host/mirror_demo"], check=True)
    s = host.sphere_splats(20000, seed=9, radius=0.02, noise=0.004)
    g, bins, lo, spacing = host.prepare(s, 0.04)
    b = bins[0]
    splats = np.ascontiguousarray(g[b["ids"]])
    path = tmp_path / "splats.bin"
    splats.tofile(path)
    args = [demo, str(path)] + [str(int(v)) for v in b["size"]] + [str(int(v)) for v in b["offset"]] + ["127"]
    out = subprocess.run(args, check=True, capture_output=True, text=True).stdout.split()
    with m.Worker(max_splats=len(splats), max_cells=127) as w:
        w.upload_splats(splats)
        blocks = w.process_bin(0, len(splats), b["size"], b["offset"])
    blk = blocks[0]
    blob = blk["external_keys"].tobytes() + blk["vertices"].tobytes() + blk["triangles"].tobytes()
    assert int(out[0]) == len(blk["vertices"]) and int(out[1]) == len(blk["triangles"])
    assert int(out[2]) == blk["num_internal"]
    assert int(out[3], 16) == fnv1a(blob)
