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
g, bins, lo, spacing = host.prepare(s, 0.012, max_cells=127)
h = hashlib.sha256()
blocks = 0
big = max(len(b["ids"]) for b in bins)
# a small mesh memory as well, so that some bins ship as several blocks (zFirst > 0 in the wire format)
for mesh_bytes in (64 << 20, 127 * 127 * 872):
    with m.Worker(max_splats=big, max_cells=127, mesh_bytes=mesh_bytes) as w:
        w.set_scale_bias(spacing, [float(v) * spacing for v in lo])
        for b in bins:
            splats = np.ascontiguousarray(g[b["ids"]])
            w.upload_splats(splats)
            for blk in w.process_bin(0, len(splats), b["size"], b["offset"]):
                blocks += 1
                for k in ("vertices", "external_keys", "triangles"):
                    h.update(np.ascontiguousarray(blk[k]).tobytes())
                h.update(str(blk["num_internal"]).encode())
print(blocks, len(bins), h.hexdigest())
""" % ROOT


def run(mode):
    env = dict(os.environ)
    env.pop("MLSB_READBACK", None)
    if mode:
        env["MLSB_READBACK"] = mode
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    blocks, bins, digest = r.stdout.split()
    return int(blocks), int(bins), digest


def test_readback_modes_give_the_same_bytes():
    base = run(None)
    assert base[0] > 2 * base[1] > 8            # several bins; the small mesh memory split them into more blocks
    assert run("pack") == base
    assert run("compact") == base
