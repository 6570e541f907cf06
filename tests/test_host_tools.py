"""Host-side tools around the device path: PLY I/O, host mesher (seam weld + prune)
and the mlsgpu command line with the reference's option names."""
import os
import subprocess

import numpy as np
import pytest

from fixtures import is_manifold
from mlsgpu_b200 import host, plyfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "mlsgpu_b200", "host")


def build(target):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "mlsgpu_b200"), "host/" + target], check=True)
    return os.path.join(HOST, target)


def tetra(offset, keys=None):
    """A tetrahedron surface: 4 vertices, 4 triangles; `keys` marks its last vertices external."""
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32) + np.float32(offset)
    t = np.array([[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]], np.uint32)
    keys = np.array(keys or [], np.uint64)
    return dict(vertices=v, triangles=t, keys=keys, num_internal=4 - len(keys))


@pytest.fixture(params=["core", "spill"])
def mesher_mode(request, tmp_path, monkeypatch):
    """core: the mesh stays in memory (finish / finishChunks); spill: the out-of-core mesher (--tmp-dir): blocks go
    to a temporary file as they complete and the output is streamed from it (Mesher::enableSpill / spill / stream).
    Every mesher fixture must come out the same either way."""
    if request.param == "spill":
        monkeypatch.setenv("MLSB_MESHER_TMP", str(tmp_path))
    else:
        monkeypatch.delenv("MLSB_MESHER_TMP", raising=False)
    return request.param


def write_blocks(path, blocks, chunks=None):
    """chunks: one (x, y, z) output chunk per block -> the harness runs the mesher in --split mode."""
    with open(path, "wb") as f:
        f.write(np.uint64(len(blocks)).tobytes())
        for b in blocks:
            f.write(np.array([len(b["vertices"]), len(b["triangles"]), b["num_internal"]], np.uint64).tobytes())
            f.write(b["keys"].tobytes() + b["vertices"].tobytes() + b["triangles"].tobytes())
        if chunks is not None:
            f.write(np.asarray(chunks, np.uint32).reshape(len(blocks), 3).tobytes())


def test_mesher_welds_by_key_and_prunes(tmp_path, mesher_mode):
    exe = build("mesher_test")
    # two big strips sharing two external vertices (keys 7 and 9) + one isolated small component
    big1 = tetra((0, 0, 0), keys=[7, 9])
    big2 = tetra((5, 5, 5), keys=[9, 7])      # same keys, other order
    many = [tetra((10 + 3 * i, 0, 0), keys=[100 + i, 101 + i]) for i in range(30)]   # a chain welded through keys
    small = tetra((-50, 0, 0))
    blocks = [big1, big2] + many + [small]
    write_blocks(tmp_path / "blocks.bin", blocks)
    out = tmp_path / "out.ply"
    res = subprocess.run([exe, str(tmp_path / "blocks.bin"), "0.1", str(out)], check=True, capture_output=True, text=True)
    nv, nt, total, kept, threshold = (int(x) for x in res.stdout.split())
    # welded vertices: big pair 4+4-2, chain 30*4 - 29 shared... (block i shares key 101+i with block i+1)
    assert nv == 6 + (30 * 4 - 29) + 4 and nt == 4 * 33
    assert total == 3 and threshold == int(nv * 0.1) and kept == 1      # only the chain is >= 10 % of the vertices
    verts, tris, header = plyfile.read_mesh_ply(out)
    assert len(verts) == 30 * 4 - 29 and len(tris) == 30 * 4
    assert tris.max() < len(verts)
    # header: padded to a multiple of 4 bytes with the reference's trailing comment (src/fast_ply.cpp:462-479)
    assert len(header) % 4 == 0 and "comment padding:" in header and "property list uint8 uint32 vertex_indices" in header
    # prune 0 keeps everything
    res = subprocess.run([exe, str(tmp_path / "blocks.bin"), "0", str(out)], check=True, capture_output=True, text=True)
    verts, tris, _ = plyfile.read_mesh_ply(out)
    assert len(verts) == nv and len(tris) == nt


@pytest.mark.parametrize("args", [
    (60000, 1, 0.01, 63, 1 << 20, 0),        # plain grid of blocks
    (60000, 2, 0.01, 63, 1500, 0),           # small --mem-bucket-splats: overfull blocks are halved
    (60000, 3, 0.01, 255, 1 << 20, 70),      # --split: chunk boundaries inside what would be one block
    (60000, 4, 0.004, 255, 20000, 300),      # both
    (50, 5, 0.05, 7, 3, 0),                  # a handful of fat splats, tiny bins
])
def test_binning_contract_by_brute_force(args):
    r = subprocess.run([build("bucket_test")] + [str(a) for a in args], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
    bins, chunks = int(r.stdout.split()[1]), int(r.stdout.split()[3])
    assert bins >= 8 and (chunks > 1) == (args[5] > 0)


def test_cli_option_validation():
    exe = build("mlsgpu")
    def run(*args):
        return subprocess.run([exe] + list(args), capture_output=True, text=True)
    assert run("--help").returncode == 0
    for args, msg in [(["in.ply"], "--output-file"), (["-o", "o.ply"], "input file"),
                      (["-o", "o.ply", "--levels", "11", "in.ply"], "--levels"),
                      (["-o", "o.ply", "--subsampling", "2", "in.ply"], "--subsampling"),
                      (["-o", "o.ply", "--fit-prune", "1.5", "in.ply"], "--fit-prune"),
                      (["-o", "o.ply", "--fit-shape", "cube", "in.ply"], "--fit-shape"),
                      (["-o", "o.ply", "--bogus", "in.ply"], "unknown option")]:
        r = run(*args)
        assert r.returncode == 1 and msg in r.stderr, (args, r.stderr)
    r = run("-o", "o.ply", "/nonexistent/in.ply", "--mem-bucket-splats=32M", "--tmp-dir", "/tmp", "--cl-gpu")
    assert r.returncode == 1 and "cannot open" in r.stderr     # ignored options are accepted


def test_cli_rejects_bad_ply(tmp_path):
    exe = build("mlsgpu")
    p = tmp_path / "bad.ply"
    p.write_bytes(b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty float32 x\nend_header\n\0\0\0\0")
    r = subprocess.run([exe, "-o", str(tmp_path / "o.ply"), str(p)], capture_output=True, text=True)
    assert r.returncode == 1 and "is missing" in r.stderr


@pytest.mark.gpu
def test_cli_end_to_end_sphere(tmp_path):
    """PLY in -> mlsgpu -> PLY out: closed manifold sphere, islands pruned, same vertex count as the
    Python pipeline over the same bins."""
    import mlsgpu_b200 as m
    from test_oracle_marching import weld_by_key
    exe = build("mlsgpu")
    s = host.sphere_splats(60000, seed=4, radius=0.012, noise=0.002)
    island = host.sphere_splats(300, seed=5, radius=0.012, noise=0.0, R=0.05)
    island["position"] += np.float32([1.6, 0, 0])
    both = np.concatenate([s, island])
    inp = tmp_path / "in.ply"
    plyfile.write_splat_ply(inp, both, extra_property=True)
    out = tmp_path / "out.ply"
    r = subprocess.run([exe, "-o", str(out), "--fit-grid", "0.025", "--levels", "3", "--fit-prune", "0.02",
                        "--device-threads", "3", "--statistics", str(inp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    stats = host.parse_statistics(r.stdout)
    assert stats["files.splats"] == len(both) and stats["bins"] >= 8 and stats["files.scans"] == 1
    assert stats["files.bytes"] == len(both) * 29            # 7 floats + the extra one-byte property
    assert stats["components.total"]["sum"] >= 2 and stats["components.kept"]["sum"] == 1
    # the registry's other kinds: peaks (bytes), per-bin variables with a spread, the marching counters
    assert stats["mem.splats"] == len(both) * 32 and stats["mem.device.all"] > 0 and stats["mem.ring"] > 0
    assert stats["copy.splats"]["n"] == stats["bins"] and stats["copy.splats"]["stddev"] is not None
    assert stats["marching.shipouts"]["n"] == stats["bins"] and stats["marching.shipouts"]["sum"] >= 8
    assert stats["bucket.collector.bins"]["sum"] == stats["bins"] and stats["run.time"]["n"] == 1
    assert stats["components.vertices.kept"]["sum"] == stats["output.vertices"] < stats["components.vertices.total"]["sum"]
    assert list(stats)[3:] == sorted(list(stats)[3:])        # version, variant, options; then name order
    verts, tris, header = plyfile.read_mesh_ply(out)
    assert is_manifold(len(verts), tris) == ""
    # the acceptance tool of the reference (extras/plymanifold.cpp): manifold, one component (the
    # noisy samples leave pin holes where the boundary test rejects corners, so boundaries exist)
    check = subprocess.run([build("plymanifold"), str(out)], capture_output=True, text=True)
    assert check.returncode == 0 and "Components: 1\n" in check.stdout, check.stdout
    assert np.abs(np.linalg.norm(verts, axis=1) - 1.0).max() < 0.03      # only the big sphere is left
    # the same input through the Python binding, bin by bin, welded by key
    g, bins, lo, spacing = host.prepare(both, 0.025, max_cells=31)
    blocks = []
    with m.Worker(max_splats=len(g), max_cells=31) as w:
        w.set_scale_bias(float(spacing), lo.astype(np.float32) * spacing)
        for b in bins:
            w.upload_splats(np.ascontiguousarray(g[b["ids"]]))
            for blk in w.process_bin(0, len(b["ids"]), b["size"], b["offset"]):
                ni = blk["num_internal"]
                blocks.append(dict(vertices=blk["vertices"], triangles=blk["triangles"], num_internal=ni,
                                   keys=np.concatenate([np.zeros(ni, np.uint64), blk["external_keys"]])))
    pv, pt = weld_by_key(blocks)
    keep = np.linalg.norm(pv, axis=1) < 1.3                      # drop the island, as --fit-prune did
    assert keep.sum() == len(verts) and keep[pt[:, 0]].sum() == len(tris)

    # --checkpoint, then --resume without input or device: the same mesh (meshdiff), and with
    # --fit-prune 0 at resume time the island comes back
    ckpt = tmp_path / "sphere.ckpt"
    r = subprocess.run([exe, "-o", str(tmp_path / "unused.ply"), "--fit-grid", "0.025", "--levels", "3", "--device-threads", "3",
                        "--checkpoint", str(ckpt), "-q", str(inp)], capture_output=True, text=True)
    assert r.returncode == 0 and ckpt.stat().st_size > 0 and not (tmp_path / "unused.ply").exists(), r.stderr
    resumed = tmp_path / "resumed.ply"
    r = subprocess.run([exe, "-o", str(resumed), "--resume", str(ckpt), "--fit-prune", "0.02", "-q"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    d = subprocess.run([build("meshdiff"), str(out), str(resumed)], capture_output=True, text=True)
    assert d.returncode == 0 and "identical up to ordering" in d.stdout, d.stdout
    r = subprocess.run([exe, "-o", str(resumed), "--resume", str(ckpt), "--fit-prune", "0", "--statistics"], capture_output=True, text=True)
    assert r.returncode == 0 and "components.kept: 2" in r.stdout
    assert len(plyfile.read_mesh_ply(resumed)[0]) == len(pv)

    # --split: a grid of chunk files that together hold exactly the same triangles (by position);
    # vertices on chunk borders appear in both neighbours; the island is pruned globally
    base = tmp_path / "split"
    r = subprocess.run([exe, "-o", str(base), "--fit-grid", "0.025", "--levels", "3", "--fit-prune", "0.02",
                        "--device-threads", "3", "--split", "--split-size", str(760 * 40 * 40), "--statistics", str(inp)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    stats = host.parse_statistics(r.stdout)
    names = sorted(tmp_path.glob("split_*.ply"))
    assert len(names) == stats["output.files"] >= 8 and stats["components.kept"]["sum"] == 1
    assert all(len(n.name) == len("split_0000_0000_0000.ply") for n in names)

    def soup(v, t):
        p = v[t]                                                  # (n, 3, 3) corner positions
        first = np.lexsort((p[:, :, 2], p[:, :, 1], p[:, :, 0]), axis=1)[:, 0]      # rotate the smallest corner first
        p = np.stack([np.roll(tri, -k, axis=0) for tri, k in zip(p, first)])
        return np.unique(p.reshape(len(p), 9), axis=0)
    parts, nverts = [], 0
    for n in names:
        v, t, _ = plyfile.read_mesh_ply(n)
        assert len(t) > 0 and t.max() < len(v) and is_manifold(len(v), t) == ""
        parts.append(soup(v, t))
        nverts += len(v)
    whole = soup(verts, tris)
    np.testing.assert_array_equal(np.unique(np.concatenate(parts), axis=0), whole)
    assert sum(len(p) for p in parts) == len(whole) == len(tris)          # no triangle twice
    assert nverts > len(verts) and nverts == stats["output.vertices"]  # border vertices are duplicated


@pytest.mark.gpu
def test_cli_plane_fit_boundary_and_prune(tmp_path):
    """A C5-style input at reduced size: near-planar terrain with a rectangular hole and isolated
    islands, plane fit, tightened boundary limit, pruning.  The result must be an oriented manifold
    WITH boundary (outer edge + hole), lie on the terrain, and contain none of the islands."""
    exe = build("mlsgpu")
    rng = np.random.default_rng(11)
    n = 80000
    xy = rng.uniform(0.0, 2.0, (n, 2))
    hole = (np.abs(xy[:, 0] - 1.0) < 0.25) & (np.abs(xy[:, 1] - 1.2) < 0.2)
    xy = xy[~hole]
    z = 0.02 * np.sin(3 * xy[:, 0]) * np.cos(2 * xy[:, 1])
    s = np.zeros(len(xy), host.SPLAT_DTYPE)
    s["position"] = np.stack([xy[:, 0], xy[:, 1], z], axis=1).astype(np.float32)
    gx = 0.06 * np.cos(3 * xy[:, 0]) * np.cos(2 * xy[:, 1])
    gy = -0.04 * np.sin(3 * xy[:, 0]) * np.sin(2 * xy[:, 1])
    nrm = np.stack([-gx, -gy, np.ones(len(xy))], axis=1)
    s["normal"] = (nrm / np.linalg.norm(nrm, axis=1)[:, None]).astype(np.float32)
    s["radius"] = 0.012
    islands = []
    for k in range(5):
        isl = np.zeros(60, host.SPLAT_DTYPE)
        c = np.array([2.6 + 0.3 * k, 0.5, 0.0])
        isl["position"] = (c + np.c_[rng.uniform(-0.03, 0.03, (60, 2)), np.zeros(60)]).astype(np.float32)
        isl["normal"] = (0, 0, 1)
        isl["radius"] = 0.012
        islands.append(isl)
    both = np.concatenate([s] + islands)
    inp, out = tmp_path / "terrain.ply", tmp_path / "terrain_mesh.ply"
    plyfile.write_splat_ply(inp, both)
    stats_file = tmp_path / "stats.txt"
    r = subprocess.run([exe, "-o", str(out), "--fit-grid", "0.02", "--fit-shape", "plane", "--fit-boundary-limit", "0.5",
                        "--fit-prune", "0.02", "--levels", "4", "--statistics-file", str(stats_file), "--statistics-cl",
                        str(inp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == ""                                   # --statistics-file: nothing on stdout
    stats = host.parse_statistics(stats_file.read_text())
    # --statistics-cl: per-kernel device times under the reference's names, "sum : mean +/- stddev [n]"
    mls = stats["kernel.mls.processCorners.time"]
    assert mls["sum"] > 0 and mls["n"] == stats["bins"] and abs(mls["sum"] / mls["n"] - mls["mean"]) < 1e-6
    assert mls["stddev"] is not None and 0 <= mls["stddev"] < mls["mean"] * 10
    assert "kernel.octree.sort.time" in stats and "kernel.marching.generateElements.time" in stats
    # (with the tightened boundary limit the tiny islands may be rejected by the boundary test already)
    assert stats["components.total"]["sum"] >= 1 and stats["components.kept"]["sum"] == 1
    verts, tris, _ = plyfile.read_mesh_ply(out)
    assert len(verts) > 5000 and is_manifold(len(verts), tris) == ""
    check = subprocess.run([build("plymanifold"), str(out)], capture_output=True, text=True)
    assert check.returncode == 0 and "Components: 1\n" in check.stdout, check.stdout
    assert int(check.stdout.strip().rsplit(": ", 1)[1]) >= 2      # boundary loops: outer rim and the hole
    assert verts[:, 0].max() < 2.2                        # islands pruned
    assert np.abs(verts[:, 2] - 0.02 * np.sin(3 * verts[:, 0]) * np.cos(2 * verts[:, 1])).max() < 0.01
    # open surface: boundary edges exist (outer rim and the hole), and no vertex sits deep inside the hole
    edges = np.sort(np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]]), axis=1)
    _, counts = np.unique(edges, axis=0, return_counts=True)
    assert (counts == 1).sum() > 100 and counts.max() == 2
    in_hole = (np.abs(verts[:, 0] - 1.0) < 0.15) & (np.abs(verts[:, 1] - 1.2) < 0.1)
    assert not in_hole.any()


# ---- the reference's own mesher fixtures (test/test_mesher.cpp:250-340, :671-743, :745-924) ----

def blk(internal, external, keys, indices):
    v = np.array(list(internal) + list(external), np.float32).reshape(-1, 3)
    return dict(vertices=v, triangles=np.array(indices, np.uint32).reshape(-1, 3), keys=np.array(keys, np.uint64),
                num_internal=len(internal))


def check_isomorphic(expected_vertices, expected_indices, verts, tris):
    """checkIsomorphic, test/test_mesher.cpp:450-520: equal up to a permutation of vertices and
    triangles and a rotation (not reflection) of each triangle; vertices are unique."""
    ev = [tuple(v) for v in np.array(expected_vertices, np.float32).tolist()]
    av = [tuple(v) for v in np.asarray(verts).tolist()]
    assert sorted(ev) == sorted(av) and len(set(av)) == len(av)

    def canon(vs, idx):
        out = []
        for t in np.array(idx).reshape(-1, 3).tolist():
            p = [vs[i] for i in t]
            k = p.index(min(p))
            out.append(tuple(p[k:] + p[:k]))
        return sorted(out)
    assert canon(ev, expected_indices) == canon(av, tris)


def test_out_of_core_mesher_writes_the_same_bytes(tmp_path, monkeypatch):
    """The out-of-core mesher against the in-core one on a few hundred welded blocks, with and without pruning,
    one file and --split: the output files are the same bytes (same vertex order, same triangle order)."""
    rng = np.random.default_rng(11)
    blocks, chunks = [], []
    for i in range(240):
        # chains of tetrahedra welded through keys, of random length, plus loose ones (pruned at 2 %)
        keys = [1000 + i, 1001 + i] if i % 40 != 39 else []
        b = tetra((3 * i, float(rng.integers(0, 5)), 0), keys=keys)
        blocks.append(b)
        chunks.append((i // 60, 0, 0))
    exe = build("mesher_test")

    def run(mode, prune, split):
        monkeypatch.delenv("MLSB_MESHER_KEEP", raising=False)
        if mode in ("spill", "mixed"):
            monkeypatch.setenv("MLSB_MESHER_TMP", str(tmp_path))
            if mode == "mixed":       # a budget that about a third of the blocks fit in: the rest goes to the file
                monkeypatch.setenv("MLSB_MESHER_KEEP", str(80 * (4 * 12 + 4 * 4 + 4 * 12)))
        else:
            monkeypatch.delenv("MLSB_MESHER_TMP", raising=False)
        d = tmp_path / ("%s_%s_%d" % (mode, prune, split))
        d.mkdir()
        write_blocks(d / "blocks.bin", blocks, chunks if split else None)
        r = subprocess.run([exe, str(d / "blocks.bin"), prune, str(d / "out"), "1"], check=True, capture_output=True, text=True)
        files = {f.name: f.read_bytes() for f in sorted(d.iterdir()) if f.name != "blocks.bin"}
        return r.stdout.replace(str(d), ""), files

    for prune in ("0", "0.02", "0.1", "0.3"):        # 0.3: everything is pruned
        for split in (0, 1):
            core = run("core", prune, split)
            for mode in ("spill", "mixed"):
                other = run(mode, prune, split)
                assert core[0] == other[0], (mode, prune, split, core[0], other[0])
                assert sorted(core[1]) == sorted(other[1]) and (len(core[1]) >= 1 or prune == "0.3")
                for name in core[1]:
                    assert core[1][name] == other[1][name], (mode, prune, split, name)
    # nothing is left behind in the temporary directory
    assert not [f for f in tmp_path.iterdir() if f.name.startswith("mlsgpu-mesher-")]


# threads > 1: the blocks go through Mesher::prepare/commit/complete concurrently, as the CLI's sink does
THREADS = pytest.mark.parametrize("threads", [1, 3])


def run_mesher(tmp_path, blocks, prune, threads=1):
    exe = build("mesher_test")
    write_blocks(tmp_path / "blocks.bin", blocks)
    out = tmp_path / "out.ply"
    subprocess.run([exe, str(tmp_path / "blocks.bin"), repr(prune), str(out), str(threads)], check=True, capture_output=True)
    verts, tris, _ = plyfile.read_mesh_ply(out)
    return verts, tris


@THREADS
def test_mesher_reference_weld_fixture(tmp_path, threads, mesher_mode):
    b0 = blk([(0, 0, 1), (0, 0, 2), (0, 0, 3), (0, 0, 4), (0, 0, 5)], [], [], [0, 1, 3, 1, 2, 3, 3, 4, 0])
    b1 = blk([], [(1, 0, 1), (1, 0, 2), (1, 0, 3), (1, 0, 4)], [0, 0x8000000000000000, 1, 0x8000000000000001],
             [0, 1, 3, 1, 2, 3, 2, 0, 3])
    b2 = blk([(0, 1, 0), (0, 2, 0), (0, 3, 0)], [(2, 0, 1), (2, 0, 2)], [0x1234567812345678, 0x12345678],
             [0, 1, 3, 1, 4, 3, 2, 3, 4, 0, 2, 4, 0, 3, 2])
    b3 = blk([(3, 3, 3)], [(4, 5, 6), (1, 0, 2), (1, 0, 3), (2, 0, 2)], [100, 0x8000000000000000, 1, 0x12345678],
             [0, 2, 1, 1, 2, 4, 4, 2, 3])
    expected_vertices = [(0, 0, 1), (0, 0, 2), (0, 0, 3), (0, 0, 4), (0, 0, 5), (1, 0, 1), (1, 0, 2), (1, 0, 3), (1, 0, 4),
                         (0, 1, 0), (0, 2, 0), (0, 3, 0), (2, 0, 1), (2, 0, 2), (3, 3, 3), (4, 5, 6)]
    expected_indices = [0, 1, 3, 1, 2, 3, 3, 4, 0, 5, 6, 8, 6, 7, 8, 7, 5, 8, 9, 10, 12, 10, 13, 12, 11, 12, 13, 9, 11, 13,
                        9, 12, 11, 14, 6, 15, 15, 6, 13, 13, 6, 7]
    verts, tris = run_mesher(tmp_path, [b0, b1, b2, b3], 0.0, threads)
    check_isomorphic(expected_vertices, expected_indices, verts, tris)


WELD_BLOCKS = [
    blk([(0, 0, 1), (0, 0, 2), (0, 0, 3), (0, 0, 4), (0, 0, 5)], [], [], [0, 1, 3, 1, 2, 3, 3, 4, 0]),
    blk([], [(1, 0, 1), (1, 0, 2), (1, 0, 3), (1, 0, 4)], [0, 0x8000000000000000, 1, 0x8000000000000001],
        [0, 1, 3, 1, 2, 3, 2, 0, 3]),
    blk([(0, 1, 0), (0, 2, 0), (0, 3, 0)], [(2, 0, 1), (2, 0, 2)], [0x1234567812345678, 0x12345678],
        [0, 1, 3, 1, 4, 3, 2, 3, 4, 0, 2, 4, 0, 3, 2]),
    blk([(3, 3, 3)], [(4, 5, 6), (1, 0, 2), (1, 0, 3), (2, 0, 2)], [100, 0x8000000000000000, 1, 0x12345678],
        [0, 2, 1, 1, 2, 4, 4, 2, 3])]


@pytest.mark.parametrize("order", [(0, 1, 2), (2, 1, 0)])
def test_mesher_reference_simple_fixture(tmp_path, order, mesher_mode):
    """testSimple, test/test_mesher.cpp:462-544: three blocks without shared keys, in either order."""
    expected_vertices = [(0, 0, 1), (0, 0, 2), (0, 0, 3), (0, 0, 4), (0, 0, 5), (1, 0, 1), (1, 0, 2), (1, 0, 3), (1, 0, 4),
                         (0, 1, 0), (0, 2, 0), (0, 3, 0), (2, 0, 1), (2, 0, 2)]
    expected_indices = [0, 1, 3, 1, 2, 3, 3, 4, 0, 5, 6, 8, 6, 7, 8, 7, 5, 8, 9, 10, 12, 10, 13, 12, 11, 12, 13, 9, 11, 13,
                        9, 12, 11]
    verts, tris = run_mesher(tmp_path, [WELD_BLOCKS[i] for i in order], 0.0)
    check_isomorphic(expected_vertices, expected_indices, verts, tris)


def test_mesher_reference_no_internal_and_no_external_fixtures(tmp_path, mesher_mode):
    """testNoInternal / testNoExternal, test/test_mesher.cpp:546-650 (degenerate triangles included)."""
    b1 = WELD_BLOCKS[1]
    b2 = blk([], [(2, 0, 1), (2, 0, 2)], [0x1234567812345678, 0x12345678], [0, 1, 1, 0, 0, 1])
    verts, tris = run_mesher(tmp_path, [b1, b2], 0.0)
    check_isomorphic([(1, 0, 1), (1, 0, 2), (1, 0, 3), (1, 0, 4), (2, 0, 1), (2, 0, 2)],
                     [0, 1, 3, 1, 2, 3, 2, 0, 3, 4, 5, 5, 4, 4, 5], verts, tris)
    b0 = WELD_BLOCKS[0]
    b2 = blk([(0, 1, 0), (0, 2, 0), (0, 3, 0)], [], [], [0, 1, 2, 2, 1, 0])
    verts, tris = run_mesher(tmp_path, [b0, b2], 0.0)
    check_isomorphic([(0, 0, 1), (0, 0, 2), (0, 0, 3), (0, 0, 4), (0, 0, 5), (0, 1, 0), (0, 2, 0), (0, 3, 0)],
                     [0, 1, 3, 1, 2, 3, 3, 4, 0, 5, 6, 7, 7, 6, 5], verts, tris)


@THREADS
def test_mesher_on_oracle_blocks(tmp_path, threads, mesher_mode):
    """A real surface shipped out in four blocks (the oracle's marching over four z ranges of the
    reference's truncated-sphere volume): the mesher's weld equals the key weld of the test suite and
    `plymanifold` accepts it — the role of testRandom, test/test_mesher.cpp:1010-1100."""
    from oracle.oracle import Oracle
    from test_oracle_marching import MARCHING_CASES, to_image, weld_by_key
    orc = Oracle()
    vol = MARCHING_CASES["truncated"]()
    d, h, w = vol.shape
    img, zs = to_image(vol)
    cuts = [0, 13, 14, 40, d - 1]                        # one block is a single slab of cells
    parts = [orc.marching_generate(img, w, h, zs, 0, z0, z1, (2, 3, 4)) for z0, z1 in zip(cuts[:-1], cuts[1:])]
    blocks = [dict(vertices=p["vertices"], triangles=p["triangles"], num_internal=p["num_internal"],
                   keys=p["keys"][p["num_internal"]:].copy()) for p in parts]
    verts, tris = run_mesher(tmp_path, blocks, 0.0, threads)
    pv, pt = weld_by_key(parts)
    whole = orc.marching_generate(img, w, h, zs, 0, 0, d - 1, (2, 3, 4))
    assert len(verts) == len(pv) == len(whole["vertices"]) and len(tris) == len(pt) == len(whole["triangles"])
    np.testing.assert_array_equal(np.unique(verts, axis=0), np.unique(whole["vertices"], axis=0))
    check = subprocess.run([build("plymanifold"), str(tmp_path / "out.ply")], capture_output=True, text=True)
    assert check.returncode == 0 and "Components: 1\n" in check.stdout, check.stdout


def run_split(tmp_path, blocks, chunks, prune, threads=1, base="chunk"):
    exe = build("mesher_test")
    write_blocks(tmp_path / "blocks.bin", blocks, chunks)
    res = subprocess.run([exe, str(tmp_path / "blocks.bin"), repr(prune), str(tmp_path / base), str(threads)],
                         check=True, capture_output=True, text=True)
    lines = res.stdout.strip().splitlines()
    files = {}
    for line in lines[1:]:
        name, nv, nt = line.rsplit(" ", 2)
        verts, tris, _ = plyfile.read_mesh_ply(name)
        assert len(verts) == int(nv) and len(tris) == int(nt)
        files[os.path.basename(name)] = (verts, tris)
    return [int(x) for x in lines[0].split()], files


@THREADS
def test_mesher_reference_chunk_fixture(tmp_path, threads, mesher_mode):
    """testChunk, test/test_mesher.cpp:924-993: the blocks of the weld fixture, each in a chunk of its
    own: every file holds its block's vertices, external ones included (no welding across files)."""
    chunks = [(i, i * i, 1) for i in range(4)]
    _, files = run_split(tmp_path, WELD_BLOCKS, chunks, 0.0, threads)
    assert sorted(files) == ["chunk_0000_0000_0001.ply", "chunk_0001_0001_0001.ply", "chunk_0002_0004_0001.ply",
                             "chunk_0003_0009_0001.ply"]
    for b, (x, y, z) in zip(WELD_BLOCKS, chunks):
        verts, tris = files["chunk_%04d_%04d_%04d.ply" % (x, y, z)]
        check_isomorphic(b["vertices"], b["triangles"].ravel(), verts, tris)


def test_mesher_split_names_and_global_components(tmp_path, mesher_mode):
    """ChunkNamer (testSimple / testBig, test/test_mesher.cpp:101-122) and what --split keeps global:
    components and pruning span the chunk files, vertices do not."""
    # a chain of tetrahedra welded through keys, laid over three chunks, plus a small island in a fourth
    chain = [tetra((3 * i, 0, 0), keys=[100 + i, 101 + i]) for i in range(12)]
    island = tetra((-50, 0, 0))
    chunks = [(0, 5, 3000)] * 4 + [(100, 123456, 2345678)] * 4 + [(1, 1, 1)] * 4 + [(7, 7, 7)]
    for prune, kept in ((0.0, 2), (0.2, 1)):
        (nv, nt, total, kept_components, threshold), files = run_split(tmp_path, chain + [island], chunks, prune, 2, base="foo")
        # 4 tetrahedra per chunk share 3 keys inside the chunk; the keys on chunk borders stay in both files
        assert nv == 3 * (16 - 3) + 4 and nt == 13 * 4
        # ... but count once for the components: 12 * 4 - 11 chained vertices + 4
        assert total == 2 and kept_components == kept and threshold == int((12 * 4 - 11 + 4) * prune)
        expected = ["foo_0000_0005_3000.ply", "foo_0001_0001_0001.ply", "foo_0100_123456_2345678.ply"]
        if kept == 2:
            expected.append("foo_0007_0007_0007.ply")       # the island's file exists only if it survives
        assert sorted(files) == sorted(expected)
        for name in expected[:3]:
            verts, tris = files[name]
            assert len(verts) == 13 and len(tris) == 16 and tris.max() < len(verts)
        for f in tmp_path.glob("foo_*.ply"):
            f.unlink()


def test_checkpoint_and_resume(tmp_path):
    """--checkpoint / --resume (OOCMesher::checkpoint / resume, src/mesher.cpp:854-898): the mesher's state
    survives a save/load round trip, and `mlsgpu --resume` writes the output from it without input,
    device or bins — with the pruning threshold of the resuming command line."""
    exe, cli = build("mesher_test"), build("mlsgpu")
    chain = [tetra((3 * i, 0, 0), keys=[100 + i, 101 + i]) for i in range(12)]
    blocks = chain + [tetra((-50, 0, 0))]
    write_blocks(tmp_path / "blocks.bin", blocks)
    ckpt = tmp_path / "state.ckpt"
    direct, resumed = tmp_path / "direct.ply", tmp_path / "resumed.ply"
    subprocess.run([exe, str(tmp_path / "blocks.bin"), "0.2", str(direct)], check=True, capture_output=True)
    subprocess.run([exe, str(tmp_path / "blocks.bin"), "0.2", str(resumed), "1", str(ckpt)], check=True, capture_output=True)
    dv, dt, _ = plyfile.read_mesh_ply(direct)
    rv, rt, _ = plyfile.read_mesh_ply(resumed)
    check_isomorphic(dv, dt.ravel(), rv, rt)
    assert len(dv) == 12 * 4 - 11                                     # the island is pruned
    # the command line resumes from the same file; --fit-prune is the resuming run's
    for prune, nv, kept in (("0.2", 12 * 4 - 11, 1), ("0", 12 * 4 - 11 + 4, 2)):
        out = tmp_path / "cli.ply"
        r = subprocess.run([cli, "-o", str(out), "--resume", str(ckpt), "--fit-prune", prune, "--statistics"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        stats = host.parse_statistics(r.stdout)
        assert stats["components.total"]["sum"] == 2 and stats["components.kept"]["sum"] == kept
        v, t, _ = plyfile.read_mesh_ply(out)
        assert len(v) == nv == stats["output.vertices"] and t.max() < len(v)
    # a --split checkpoint resumes into chunk files
    chunks = [(0, 0, 0)] * 6 + [(1, 0, 0)] * 6 + [(2, 2, 2)]
    write_blocks(tmp_path / "blocks.bin", blocks, chunks)
    subprocess.run([exe, str(tmp_path / "blocks.bin"), "0", str(tmp_path / "h"), "1", str(ckpt)], check=True, capture_output=True)
    r = subprocess.run([cli, "-o", str(tmp_path / "r"), "--resume", str(ckpt), "--fit-prune", "0", "-q"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == ""
    for name in ("_0000_0000_0000.ply", "_0001_0000_0000.ply", "_0002_0002_0002.ply"):
        hv, ht, _ = plyfile.read_mesh_ply(str(tmp_path / "h") + name)
        cv, ct, _ = plyfile.read_mesh_ply(str(tmp_path / "r") + name)
        check_isomorphic(hv, ht.ravel(), cv, ct)
    # garbage is refused
    (tmp_path / "bad.ckpt").write_bytes(b"not a checkpoint at all")
    r = subprocess.run([cli, "-o", str(tmp_path / "x.ply"), "--resume", str(tmp_path / "bad.ckpt")], capture_output=True, text=True)
    assert r.returncode == 1 and "checkpoint" in r.stderr


@THREADS
def test_mesher_reference_prune_fixture(tmp_path, threads, mesher_mode):
    b0 = blk([(0, 0, 0), (1, 0, 0), (2, 0, 0), (3, 0, 0), (4, 0, 0)], [(0, 3, 0), (1, 3, 0), (2, 3, 0)], [0x30, 0x31, 0x32],
             [0, 4, 1, 1, 4, 2, 2, 4, 3, 5, 7, 6])
    b1 = blk([(0, 1, 0), (1, 1, 0), (2, 1, 0), (3, 1, 0), (4, 1, 0), (5, 1, 0), (0, 2, 0), (3, 2, 0)],
             [(2, 2, 0), (4, 2, 0), (0, 3, 0), (2, 3, 0), (4, 3, 0)], [0x22, 0x24, 0x30, 0x32, 0x34],
             [0, 5, 1, 1, 5, 2, 2, 5, 3, 3, 5, 4, 6, 7, 9, 9, 7, 8, 10, 12, 11])
    b2 = blk([], [(1, 3, 0), (2, 3, 0), (3, 3, 0)], [0x31, 0x32, 0x33], [0, 1, 2])
    b3 = blk([(1, 2, 0), (5, 3, 0)], [(2, 2, 0), (3, 3, 0), (4, 2, 0), (4, 3, 0), (2, 3, 0)], [0x22, 0x33, 0x24, 0x34, 0x32],
             [6, 5, 3, 4, 2, 0, 3, 5, 1])
    expected_vertices = [(0, 1, 0), (1, 1, 0), (2, 1, 0), (3, 1, 0), (4, 1, 0), (5, 1, 0), (0, 3, 0), (1, 3, 0), (2, 3, 0),
                         (3, 3, 0), (4, 3, 0), (5, 3, 0)]
    expected_indices = [0, 5, 1, 1, 5, 2, 2, 5, 3, 3, 5, 4, 6, 8, 7, 7, 8, 9, 9, 8, 10, 9, 10, 11, 6, 10, 8]
    # 22 vertices in total, threshold 6 (setPruneThreshold(6.5 / 22.0))
    verts, tris = run_mesher(tmp_path, [b0, b1, b2, b3], 6.5 / 22.0, threads)
    check_isomorphic(expected_vertices, expected_indices, verts, tris)


@THREADS
def test_mesher_reference_empty_fixture(tmp_path, threads, mesher_mode):
    # TestMesherBase::testEmpty: no blocks (or empty blocks) give an empty, well-formed PLY
    verts, tris = run_mesher(tmp_path, [], 0.0, threads)
    assert len(verts) == 0 and len(tris) == 0
    verts, tris = run_mesher(tmp_path, [blk([], [], [], [])], 0.02, threads)
    assert len(verts) == 0 and len(tris) == 0


# ---- out-of-core front-end (stream.hpp): BucketCollector / BucketLoader over chunked, re-read input ----

def synth(tmp_path, config, points):
    exe = build("gen_synth")
    r = subprocess.run([exe, "--config", config, "--points", str(points), "--scans", "4", "--out", str(tmp_path / "in")],
                       check=True, capture_output=True, text=True)
    fit_grid = float(r.stdout.split("fit-grid ")[1])
    return sorted(str(p) for p in tmp_path.glob("in_*.ply")), fit_grid


@pytest.mark.parametrize("config,points,max_cells,max_bin,max_load,chunk_cells", [
    ("C3", 600_000, 255, 1 << 21, 60_000, 0),       # many batches of whole blocks
    ("C3", 600_000, 63, 5_000, 20_000, 150),        # blocks halved by the splat limit, --split chunks
    ("C5", 400_000, 127, 1 << 21, 1 << 30, 0),      # one batch: everything at once
])
def test_streamed_bins_equal_in_core_bins(tmp_path, config, points, max_cells, max_bin, max_load, chunk_cells):
    """The streaming front-end must hand over, for every bin, the same box, chunk and splats in the same
    order as binning the whole input in memory — while never loading more than the limit at once."""
    files, fit_grid = synth(tmp_path, config, points)
    r = subprocess.run([build("stream_test"), repr(fit_grid), str(max_cells), str(max_bin), str(max_load), str(chunk_cells)] + files,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith("ok:"), r.stdout
    batches = int(r.stdout.split(" bins, ")[1].split(" batches")[0])
    assert (batches > 3) == (max_load < points)


def test_statistics_registry_known_answers():
    """host/statistics.hpp against the reference's own cases (test/test_statistics.cpp): counters, variables
    (mean, sample stddev, length_error below two samples), peaks, registry order / merge / type clash, threads."""
    r = subprocess.run([build("statistics_test")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "statistics_test: ok", r.stdout + r.stderr


def test_compact_readback_unpack_known_answers():
    """csrc/unpack.cpp (host half of MLSB_READBACK=compact) against a scalar restatement of the device's vertex
    arithmetic on random wire blobs: bit for bit, subnormal / -0 / huge t included, nothing written past the blob."""
    r = subprocess.run([build("unpack_test")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("unpack_test: ok"), r.stdout + r.stderr


def test_parse_statistics_formats():
    text = ("mlsgpu version: x\ncounter: 100\nstat0: [0]\nstat1: 1 : 1 [1]\nstat3: 12 : 4 +/- 2 [3]\n"
            "splats/s end-to-end: 2.5e+07 : 2.5e+07 [1]\n")
    st = host.parse_statistics(text)
    assert st["mlsgpu version"] == "x" and st["counter"] == 100
    assert st["stat0"] == {"n": 0, "sum": 0.0, "mean": None, "stddev": None}
    assert st["stat1"] == {"n": 1, "sum": 1.0, "mean": 1.0, "stddev": None}
    assert st["stat3"] == {"n": 3, "sum": 12.0, "mean": 4.0, "stddev": 2.0}
    assert st["splats/s end-to-end"]["mean"] == 2.5e7


@pytest.mark.gpu
def test_cli_streamed_run_equals_in_core_run(tmp_path):
    """mlsgpu with the reference's kind of --mem-load-splats (a fraction of the input): several passes
    over the files, bounded host memory for splats, and the same mesh as the in-core run."""
    files, fit_grid = synth(tmp_path, "C3", 1_500_000)
    exe = build("mlsgpu")
    common = [exe, "--fit-grid", repr(fit_grid), "--device", "0", "--device-threads", "3", "--statistics"]
    a = subprocess.run(common + ["-o", str(tmp_path / "core.ply"), "--in-core"] + files, capture_output=True, text=True)
    b = subprocess.run(common + ["-o", str(tmp_path / "stream.ply"), "--mem-load-splats", "8M", "--mem-bucket-splats", "4M"] + files,
                       capture_output=True, text=True)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    assert "stream.batches" not in a.stdout
    sb = host.parse_statistics(b.stdout)
    assert sb["stream.batches"] >= 4 and sb["files.splats"] == 1_500_000
    # one load / compute sample per batch; the splats held at once stayed near the limit (8 MB = 262 144 splats;
    # a batch is a run of whole blocks, so it may exceed it by one block)
    assert sb["bucket.loader.load"]["n"] == sb["stream.batches"] == sb["bucket.loader.compute"]["n"]
    assert sb["mem.splats"] < host.parse_statistics(a.stdout)["mem.splats"] / 2
    d = subprocess.run([build("meshdiff"), str(tmp_path / "core.ply"), str(tmp_path / "stream.ply")], capture_output=True, text=True)
    assert d.returncode == 0, d.stdout + d.stderr
    # --in-core never opens a temporary file; by default there is one, unused while the mesh fits --mem-host-mesh
    assert "mesher.tmp.bytes" not in a.stdout and sb["mesher.tmp.bytes"] == 0
    # out of core at both ends: streamed input AND the mesher spilling all but 4 MB of finished blocks to --tmp-dir
    spill_dir = tmp_path / "tmp"
    spill_dir.mkdir()
    c = subprocess.run(common + ["-o", str(tmp_path / "ooc.ply"), "--mem-load-splats", "8M", "--mem-bucket-splats", "4M",
                                 "--tmp-dir", str(spill_dir), "--mem-host-mesh", "4M"] + files, capture_output=True, text=True)
    assert c.returncode == 0, c.stderr
    sc = host.parse_statistics(c.stdout)
    assert sc["mesher.tmp.bytes"] >= 28 * sc["output.vertices"] - (6 << 20) and list(spill_dir.iterdir()) == []
    assert sc["output.vertices"] == sb["output.vertices"] and sc["output.triangles"] == sb["output.triangles"]
    d = subprocess.run([build("meshdiff"), str(tmp_path / "core.ply"), str(tmp_path / "ooc.ply")], capture_output=True, text=True)
    assert d.returncode == 0, d.stdout + d.stderr


@pytest.mark.gpu
def test_cli_timeplot(tmp_path):
    """--timeplot: EVENT lines in the reference's format (src/timeplot.cpp:134-186) for the loader, every
    device worker thread (get / copy / compute per bin) and the mesher, on one clock."""
    files, fit_grid = synth(tmp_path, "C3", 400_000)
    log = tmp_path / "timeplot.txt"
    r = subprocess.run([build("mlsgpu"), "-o", str(tmp_path / "o.ply"), "--fit-grid", repr(fit_grid), "--device", "0",
                        "--device-threads", "2", "--timeplot", str(log), "--statistics"] + files, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    bins = host.parse_statistics(r.stdout)["bins"]
    events = [l.split() for l in log.read_text().splitlines()]
    assert all(e[0] == "EVENT" and len(e) == 5 and float(e[3]) <= float(e[4]) for e in events)
    workers = {e[1] for e in events}
    assert {"main", "mesher", "device.0.0", "device.0.1"} <= workers
    compute = [e for e in events if e[1].startswith("device.") and e[2] == "compute"]
    assert len(compute) == bins and len([e for e in events if e[2] == "copy"]) == bins
    load = [e for e in events if e[1] == "main" and e[2] == "load"][0]
    assert float(load[4]) <= min(float(e[3]) for e in compute) + 1e-3      # bins are computed after they were loaded
