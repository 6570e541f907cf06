#!/usr/bin/env python
"""Generates one of the large synthetic configurations of SURVEY 8(d) (C3: multi-scan terrain over
several GPUs; C5: near-planar terrain with holes, ragged rim and islands), runs the `mlsgpu` command
line on it — ONE process, one worker group per device, one host mesher welding the seams — checks the
output with `plymanifold` and prints one JSON line.

    python tools/run_config.py --config C3 --points 100000000 --devices all
    python tools/run_config.py --config C5 --points 50000000 --boundary-limit 0.5
    python tools/run_config.py --config C3 --label C4 --points 1000000000 --scans 80 --fit-grid 0.0011 --devices 0
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "mlsgpu_b200", "host")


sys.path.insert(0, ROOT)


def stats_dict(text):
    """The CLI's statistics, flattened: counters as they are, variables by their sum (one sample each for the
    run-level times), as strings so that the JSON line reads like the listing."""
    from mlsgpu_b200.host import parse_statistics
    out = {}
    for k, v in parse_statistics(text).items():
        out[k] = ("%.6g" % v["sum"]) if isinstance(v, dict) else str(v)
    return out


def run(config, points, devices="all", threads=4, scans=8, boundary_limit=1.0, prune=0.02, shape="sphere", keep=None, extra=(),
        tmp_root=None, check=True, grid=None, label=None):
    with tempfile.TemporaryDirectory(dir=tmp_root) as tmp:
        t0 = time.time()
        g = subprocess.run([os.path.join(HOST, "gen_synth"), "--config", config, "--points", str(points), "--scans", str(scans),
                            "--out", os.path.join(tmp, "in")], capture_output=True, text=True, check=True)
        m = re.search(r"points (\d+) .*fit-grid ([0-9.e+-]+)", g.stdout)
        n, fit_grid = int(m.group(1)), float(m.group(2))
        if grid:
            fit_grid = grid
        t_gen = time.time() - t0
        inputs = sorted(os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".ply"))
        out = os.path.join(tmp, "mesh.ply")
        cmd = [os.path.join(HOST, "mlsgpu"), "-o", out, "--fit-grid", repr(fit_grid), "--fit-prune", repr(prune),
               "--fit-boundary-limit", repr(boundary_limit), "--fit-shape", shape, "--device-threads", str(threads),
               "--device", devices, "--statistics"] + list(extra) + inputs
        t0 = time.time()
        r = subprocess.run(cmd, capture_output=True, text=True)
        wall = time.time() - t0
        res = {"config": label or config, "points": n, "scans": scans, "fit_grid": fit_grid, "devices": devices, "device_threads": threads,
               "boundary_limit": boundary_limit, "prune": prune, "generate_s": round(t_gen, 2), "mlsgpu_wall_s": round(wall, 2),
               "exit": r.returncode}
        if r.returncode != 0:
            res["stderr"] = r.stderr[-2000:]
            return res
        st = stats_dict(r.stdout)
        res["statistics"] = {k: st[k] for k in st if k.split(".")[0] in ("files", "bins", "devices", "components", "output", "read", "bucket", "stream", "mem", "marching",
                                                                        "device", "mesher", "run") or k.startswith("splats/s")}
        res["Msplats_per_s_end_to_end"] = n / float(st["run.time"]) / 1e6
        if check:
            t0 = time.time()
            c = subprocess.run([os.path.join(HOST, "plymanifold"), out], capture_output=True, text=True)
            res["plymanifold"] = c.stdout.strip().replace("\n", "; ")
            res["plymanifold_s"] = round(time.time() - t0, 2)
        res["output_bytes"] = os.path.getsize(out)
        if keep:
            os.replace(out, keep)
        return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3", choices=["C3", "C5"])
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--devices", default="all")
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--scans", type=int, default=8)
    ap.add_argument("--boundary-limit", type=float, default=1.0)
    ap.add_argument("--prune", type=float, default=0.02)
    ap.add_argument("--shape", default="sphere")
    ap.add_argument("--keep", default=None, help="move the output mesh here")
    ap.add_argument("--tmp", default=None)
    ap.add_argument("--fit-grid", type=float, default=None, help="override the generator's mean sample spacing")
    ap.add_argument("--label", default=None, help="name of the configuration in the JSON line (C4 = C3's generator, 80 scans, streamed)")
    ap.add_argument("--no-check", action="store_true", help="skip plymanifold (minutes for a 100 M-splat mesh)")
    ap.add_argument("extra", nargs="*")
    a = ap.parse_args()
    res = run(a.config, a.points, a.devices, a.threads, a.scans, a.boundary_limit, a.prune, a.shape, a.keep, a.extra, a.tmp, not a.no_check, a.fit_grid, a.label)
    print(json.dumps(res), flush=True)
    return 0 if res["exit"] == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
