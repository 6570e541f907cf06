#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root).

The reference itself cannot be built or run in this environment (no Boost / clogs / OpenCL /
cppunit), so these vectors are outputs of the oracle — which is pinned by the reference's own
known-answer tests (tests/test_oracle_*.py) — frozen on the build machine.  They guard against
the oracle drifting (compiler flags, a different host CPU on the GPU box) and give the CUDA path
a second, file-based anchor.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.oracle import Oracle  # noqa: E402
from test_oracle_mls import process_corners_case  # noqa: E402
from test_oracle_marching import sphere_field, to_image  # noqa: E402
from test_oracle_octree import random_splats  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    orc = Oracle()
    # 1. processCorners on the reference's fixture (test/test_mls.cpp:416-514)
    c = process_corners_case(orc, 35.0)
    orc.mls_process_corners(c["image"], c["splats"], c["commands"], c["start"], 3 * c["sub"], c["offset"],
                            c["size"][0], c["size"][1], c["z_stride"], c["z_bias"], c["z_first"], c["z_last"],
                            orc.boundary_factor(1.0), 0)
    rows = slice(c["z_first"] * c["z_stride"] + c["z_bias"], (c["z_last"] + 1) * c["z_stride"] + c["z_bias"])
    np.savez_compressed(os.path.join(HERE, "mls_process_corners.npz"), field=c["image"][rows].view(np.uint32))
    # 2. octree of the reference's random fixture (test/test_splat_tree.cpp:172-200)
    splats, cells = random_splats()
    t = orc.octree_build(splats, 0, len(splats), (1, 2, -1), subsampling=2, max_levels=8)
    np.savez_compressed(os.path.join(HERE, "octree_random.npz"), start=t["start"], commands=t["commands"][:t["used"]],
                        inv_r2=splats["radius"].view(np.uint32))
    # 3. marching: a sphere truncated by the box, key offset, in grid coordinates
    vol = sphere_field(40, 36, 33, 19.5, 17.0, 16.0, 21.0)
    img, zs = to_image(vol)
    m = orc.marching_generate(img, 40, 36, zs, 0, 0, 32, (7, 100, 3))
    np.savez_compressed(os.path.join(HERE, "marching_sphere.npz"), vertices=m["vertices"].view(np.uint32),
                        keys=m["keys"][m["num_internal"]:], triangles=m["triangles"],
                        num_internal=np.array([m["num_internal"]]))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
