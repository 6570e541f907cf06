#!/usr/bin/env python
"""End-to-end run of the mlsgpu command line on the C2 workload: writes the synthetic torus as a
binary PLY, runs `mlsgpu`, checks the output and prints the tool's own statistics
(splats/s end-to-end = input splats / wall time of read + bucket + device + mesher + write)."""
import argparse
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--splats", type=int, default=10_000_000)
    ap.add_argument("--fit-grid", type=float, default=0.002)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--check", action="store_true", help="run plymanifold on the output")
    args = ap.parse_args()
    from mlsgpu_b200 import host, plyfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "mlsgpu_b200", "host", "mlsgpu")
    s = host.torus_splats(args.splats, seed=2)
    with tempfile.TemporaryDirectory() as tmp:
        inp, out = os.path.join(tmp, "torus.ply"), os.path.join(tmp, "mesh.ply")
        plyfile.write_splat_ply(inp, s)
        for rep in range(2):
            t0 = time.time()
            r = subprocess.run([exe, "-o", out, "--fit-grid", str(args.fit_grid), "--device-threads", str(args.threads),
                                "--statistics", inp], capture_output=True, text=True)
            print("run %d: exit %d, %.2f s wall" % (rep, r.returncode, time.time() - t0))
            print(r.stdout.strip())
            if r.returncode != 0:
                print(r.stderr)
                return 1
        verts, tris, header = plyfile.read_mesh_ply(out)
        print("output: %d vertices, %d triangles, %.1f MB" % (len(verts), len(tris), os.path.getsize(out) / 1e6))
        if args.check:
            t0 = time.time()
            r = subprocess.run([os.path.join(root, "mlsgpu_b200", "host", "plymanifold"), out], capture_output=True, text=True)
            print("plymanifold (%.1f s): %s" % (time.time() - t0, r.stdout.strip().replace("\n", "; ")))
    return 0


if __name__ == "__main__":
    sys.exit(main())
