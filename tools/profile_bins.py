#!/usr/bin/env python
"""Runs a few bins of the bench workload through one worker — a short command to wrap in ncu.

    ncu --set full --clock-control none --import-source on -k regex:processCorners -s 2 -c 3 \
        -o gpurun_out/prof python tools/profile_bins.py --bins 6
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--splats", type=int, default=10_000_000)
    ap.add_argument("--bins", type=int, default=6)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--stats", action="store_true")
    args = ap.parse_args()
    import bench
    import mlsgpu_b200 as m
    wl = bench.make_workload(args.splats, 2)
    bins = sorted(wl["bins"], key=lambda b: -b["n"])[:args.bins]
    with m.Worker(max_splats=max(b["n"] for b in bins), max_cells=255, mesh_bytes=256 << 20) as w:
        w.set_scale_bias(float(wl["spacing"]), wl["lo"].astype(np.float32) * wl["spacing"])
        for it in range(args.repeat + (1 if args.stats else 0)):
            if args.stats and it == 1:      # the first pass is a warm-up (module load, first-touch allocations)
                w.set_profiling(2)
                w.reset_statistics()
            for b in bins:
                s = np.ascontiguousarray(wl["splats"][b["first"]:b["first"] + b["n"]])
                w.upload_splats(s)
                blocks = w.process_bin(0, b["n"], b["size"], b["offset"])
                if not args.stats:
                  print("bin n=%d size=%s -> %d vertices, %d triangles" % (
                    b["n"], tuple(int(v) for v in b["size"]), sum(len(x["vertices"]) for x in blocks),
                    sum(len(x["triangles"]) for x in blocks)), flush=True)
        if args.stats:
            for name in bench.OCTREE_STATS + [bench.MLS_STAT] + bench.MARCHING_STATS:
                s_, c_ = w.statistic(name)
                print("  %-44s %9.3f ms  %5d launches" % (name, 1e3 * s_, c_))


if __name__ == "__main__":
    main()
