"""The large synthetic configurations at reduced scale through the `mlsgpu` command line (one
process, one worker group per device, one host mesher): C3 (multi-scan terrain, seams between bins
and between DEVICES welded by key) and C5 (near-planar terrain, holes, ragged rim, islands; boundary
limit and pruning).  tools/run_config.py runs the same at full size (profiles/)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "mlsgpu_b200"), "host/gen_synth", "host/mlsgpu", "host/plymanifold",
                    "host/meshdiff"], check=True)


def device_count():
    import mlsgpu_b200 as m
    return m.device_count()


def test_c3_multi_scan_terrain_single_process(tmp_path):
    """2 M points in 8 overlapping scans: tens of bins, every device of the box in one process; the
    result is a manifold with one component, and it is the same mesh whatever the number of devices."""
    import run_config
    build()
    one = run_config.run("C3", 2_000_000, devices="0", threads=3, keep=str(tmp_path / "one.ply"))
    assert one["exit"] == 0, one
    st = one["statistics"]
    assert int(st["bins"]) >= 16 and int(st["files.splats"]) == 2_000_000
    assert one["plymanifold"].startswith("Mesh is manifold."), one["plymanifold"]
    assert int(st["components.kept"]) == 1
    n = device_count()
    if n >= 2:
        allg = run_config.run("C3", 2_000_000, devices="all", threads=3, keep=str(tmp_path / "all.ply"))
        assert allg["exit"] == 0, allg
        sa = allg["statistics"]
        assert int(sa["devices"]) == n
        # every device got bins, and the bins met at their seams: same vertex / triangle counts, same mesh
        assert all(int(sa["device.%d.bins" % d]) > 0 for d in range(n))
        assert sa["output.vertices"] == st["output.vertices"] and sa["output.triangles"] == st["output.triangles"]
        d = subprocess.run([os.path.join(ROOT, "mlsgpu_b200", "host", "meshdiff"), str(tmp_path / "one.ply"), str(tmp_path / "all.ply")],
                           capture_output=True, text=True)
        assert d.returncode == 0, d.stdout + d.stderr


@pytest.mark.parametrize("limit", [1.0, 0.5])
def test_c5_planar_terrain_boundary_and_prune(tmp_path, limit):
    """1.5 M near-planar points with 64 holes, a ragged rim and 200 islands of 500 points: the sheet
    comes out as ONE component with boundaries (the holes and the rim), every island is pruned, and a
    tighter boundary limit cuts the mesh back from the holes (fewer triangles)."""
    import run_config
    build()
    res = run_config.run("C5", 1_500_000, devices="0", threads=3, boundary_limit=limit, prune=0.02)
    assert res["exit"] == 0, res
    st = res["statistics"]
    assert int(st["components.total"]) >= 150, st          # the islands were meshed ...
    assert int(st["components.kept"]) == 1, st             # ... and pruned: only the sheet is left
    assert int(st["output.triangles"]) > 1_000_000
    assert res["plymanifold"].startswith("Mesh is manifold.") and "Components: 1;" in res["plymanifold"], res["plymanifold"]
    assert int(res["plymanifold"].split("Boundaries: ")[1]) >= 60      # the holes and the rim are open boundaries
    test_c5_planar_terrain_boundary_and_prune.triangles[limit] = int(st["output.triangles"])
    t = test_c5_planar_terrain_boundary_and_prune.triangles
    if len(t) == 2:
        assert t[0.5] < t[1.0]


test_c5_planar_terrain_boundary_and_prune.triangles = {}
