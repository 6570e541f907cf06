#!/usr/bin/env python
"""Per-stage timing of ONE large bin (BASELINE.json configs[0]: C1, noisy sphere, single bin):
how the HBM-bound stages behave when a bin is well populated (the 70 bins of C2 hold ~150 k
splats each, which leaves octree and marching launch-latency-bound)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--splats", type=int, default=2_000_000)
    ap.add_argument("--repeat", type=int, default=5)
    args = ap.parse_args()
    import bench
    import mlsgpu_b200 as m
    from mlsgpu_b200 import host
    s = host.sphere_splats(args.splats, seed=1, radius=0.005, noise=0.002)
    g, bins, lo, spacing = host.prepare(s, 0.01, max_splats=1 << 22)
    assert len(bins) == 1, len(bins)
    b = bins[0]
    splats = np.ascontiguousarray(g[b["ids"]])
    n = len(splats)
    with m.Worker(max_splats=n, max_cells=255, mesh_bytes=512 << 20) as w:
        w.set_scale_bias(float(spacing), lo.astype(np.float32) * spacing)
        for it in range(args.repeat + 1):
            if it == 1:
                w.set_profiling(2)
                w.reset_statistics()
            w.upload_splats(splats)
            blocks = w.process_bin(0, n, b["size"], b["offset"])
        nv = sum(len(x["vertices"]) for x in blocks)
        nt = sum(len(x["triangles"]) for x in blocks)
        print("bin: %d splats, size %s -> %d vertices, %d triangles" % (n, tuple(int(v) for v in b["size"]), nv, nt))
        tot = {}
        for group, names in (("octree", bench.OCTREE_STATS), ("mls", [bench.MLS_STAT]), ("marching", bench.MARCHING_STATS)):
            tot[group] = 0.0
            for name in names:
                s_, c_ = w.statistic(name)
                tot[group] += s_ / args.repeat
                print("  %-44s %9.1f us" % (name, 1e6 * s_ / args.repeat))
        oct_bytes = 472.0 * n
        # SURVEY 8d marching model, as bench.py states it: field read 4 B per corner; per occupied cell 24 W + 56 R; per
        # per-cell vertex 24 W + a 4-pass (8 B key, 4 B index) sort read and written + 36 R; per welded vertex 24 W;
        # per index 4 W + 12 (reindex).  The counts are the worker's own (mlsb_worker_get_marching_work).
        mc, mk, mv, mi, mw = (x / args.repeat for x in w.marching_work())
        mar_bytes = 4.0 * mc + 80.0 * mk + 156.0 * mv + 24.0 * mw + 16.0 * mi
        print("octree   : %.1f us, model %.0f MB -> %.0f GB/s (%.1f%% of 6541 GB/s)" % (
            1e6 * tot["octree"], oct_bytes / 1e6, oct_bytes / tot["octree"] / 1e9, 100 * oct_bytes / tot["octree"] / 6541.1e9))
        print("marching : %.1f us, model %.0f MB -> %.0f GB/s (%.1f%% of 6541 GB/s)" % (
            1e6 * tot["marching"], mar_bytes / 1e6, mar_bytes / tot["marching"] / 1e9, 100 * mar_bytes / tot["marching"] / 6541.1e9))
        print("mls      : %.1f us" % (1e6 * tot["mls"]))


if __name__ == "__main__":
    main()
