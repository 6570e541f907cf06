"""The three ways a shipped block reaches the host (MLSB_READBACK): copy engine (default), pack
kernel writing pinned memory, compact wire format + host unpack.  The mode is fixed per process, so each runs
in a process of its own; the meshes must be the same bytes."""
import hashlib
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, %r)
import mlsgpu_b200 as m
from mlsgpu_b200 import host
s = host.torus_splats(300000, seed=5)
g, bins, lo, spacing = host.prepare(s, 0.008, max_cells=127)
h = hashlib.sha256()
blocks = 0
big = max(len(b["ids"]) for b in bins)


def add(blk):
    global blocks
    blocks += 1
    for k in ("vertices", "external_keys", "triangles"):
        h.update(np.ascontiguousarray(blk[k]).tobytes())
    h.update(str(blk["num_internal"]).encode())


with m.Worker(max_splats=big, max_cells=127) as w:
    w.set_scale_bias(spacing, [float(v) * spacing for v in lo])
    for b in bins:
        splats = np.ascontiguousarray(g[b["ids"]])
        w.upload_splats(splats)
        for blk in w.process_bin(0, len(splats), b["size"], b["offset"]):
            add(blk)
# the minimum mesh memory: one volume shipped as several blocks (zFirst > 0 in the wire format), with a key offset
sys.path.insert(0, %r)
from test_oracle_marching import MARCHING_CASES
with m.Worker(max_splats=16, max_cells=82, mesh_bytes=82 * 82 * 872) as w:
    w.set_scale_bias(0.25, [1.0, -2.0, 3.5])
    split = w.marching_generate_host_field(MARCHING_CASES["truncated"](), (5, 6, 7))
    for blk in split:
        add(blk)
print(blocks, len(bins), len(split), h.hexdigest())

""" % (ROOT, os.path.join(ROOT, "tests"))


def run(mode):
    env = dict(os.environ)
    env.pop("MLSB_READBACK", None)
    if mode:
        env["MLSB_READBACK"] = mode
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    blocks, bins, split, digest = r.stdout.split()
    return int(blocks), int(bins), int(split), digest


def test_readback_modes_give_the_same_bytes():
    base = run(None)
    assert base[0] >= base[1] >= 8 and base[2] > 1   # several bins, and a volume that ships as several blocks
    assert run("pack") == base
    assert run("compact") == base
