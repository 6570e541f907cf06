"""Host-side input conditioning and binning that sits just in front of the device path.

Mirrors what the reference does between a PLY vertex and the splat buffer of a bin:
FastPly::Reader::decode (src/fast_ply.cpp:334-350), the bounding grid of
SplatSet::FastBlobSet::getBoundingGrid (src/splat_set_impl.h:771-810), and
BucketLoader::operator() (src/bucket_loader.cpp:56-128).  The spatial partition here is a
plain recursive halving of the grid (the reference's histogram-driven Bucket::bucket,
src/bucket_impl.h:439-561, is out of scope for this round); bins have the same contract:
at most ``max_cells`` cells per side, at most ``max_splats`` splats, and every splat whose
ball reaches the bin is included.
"""
import numpy as np

from .capi import SPLAT_DTYPE


def condition_splats(splats, smooth=4.0, max_radius=np.inf):
    """radius <- min(radius, maxRadius) * smooth; quality <- 1/radius^2 (double division)."""
    r = np.minimum(splats["radius"], np.float32(max_radius)) * np.float32(smooth)
    splats["radius"] = r
    splats["quality"] = (1.0 / (r * r).astype(np.float64)).astype(np.float32)
    return splats


def bounding_grid(splats, spacing, bucket_size=63):
    """Vertex-index extents [lo, hi] per axis of the full grid (reference point = origin)."""
    pos = splats["position"].astype(np.float64)
    r = splats["radius"].astype(np.float64)[:, None]
    lo = np.floor((pos - r).min(axis=0) / spacing).astype(np.int64)
    hi = np.ceil((pos + r).max(axis=0) / spacing).astype(np.int64)
    lo = (lo // bucket_size) * bucket_size   # rounded down to a multiple of the blob bucket
    return lo, hi


def world_to_grid(splats, spacing, lo):
    """Grid::worldToVertex + radius rescale, in float32 exactly as src/grid.cpp:99-106."""
    inv = np.float32(1.0) / np.float32(spacing)
    out = splats.copy()
    out["position"] = splats["position"] * inv - lo.astype(np.float32)[None, :]
    out["radius"] = splats["radius"] * inv
    return out


def make_bins(grid_splats, num_cells, max_cells=255, max_splats=1 << 21):
    """Partition cells [0, num_cells) into bins.  Returns a list of dicts with
    offset (int32[3]), size = corners per axis (uint32[3]) and ids (indices into grid_splats,
    ascending)."""
    pos = grid_splats["position"]
    r = grid_splats["radius"]
    lo_all = pos - r[:, None]
    hi_all = pos + r[:, None]
    bins = []

    def recurse(lo, hi, ids):
        # splats whose ball reaches the closed corner box [lo, hi]
        m = np.ones(len(ids), bool)
        for a in range(3):
            m &= (hi_all[ids, a] >= lo[a]) & (lo_all[ids, a] <= hi[a])
        ids = ids[m]
        if len(ids) == 0:
            return
        ext = hi - lo
        if ext.max() <= max_cells and len(ids) <= max_splats:
            bins.append(dict(offset=lo.astype(np.int32), size=(ext + 1).astype(np.uint32), ids=ids))
            return
        a = int(np.argmax(ext))
        if ext[a] <= 1:
            raise ValueError("too many splats over a single cell (DensityError in the reference)")
        mid = lo[a] + ext[a] // 2
        hi0 = hi.copy()
        hi0[a] = mid
        lo1 = lo.copy()
        lo1[a] = mid
        recurse(lo, hi0, ids)
        recurse(lo1, hi, ids)

    # first cut the grid into max_cells-sized blocks so that bins are as cubic as possible;
    # candidate (splat, block) pairs come from each splat's bounding box in block units
    n = np.asarray(num_cells, np.int64)
    counts = -(-n // max_cells)
    edges = [np.arange(counts[a] + 1) * n[a] // counts[a] for a in range(3)]
    first = [np.clip(np.searchsorted(edges[a], lo_all[:, a], side="left") - 1, 0, counts[a] - 1) for a in range(3)]
    last = [np.clip(np.searchsorted(edges[a], hi_all[:, a], side="right") - 1, 0, counts[a] - 1) for a in range(3)]
    span = [last[a] - first[a] + 1 for a in range(3)]
    per = span[0] * span[1] * span[2]
    sid = np.repeat(np.arange(len(grid_splats)), per)
    k = np.arange(len(sid)) - np.repeat(np.cumsum(per) - per, per)
    sx, sy = span[0][sid], span[1][sid]
    bx = first[0][sid] + k % sx
    by = first[1][sid] + (k // sx) % sy
    bz = first[2][sid] + k // (sx * sy)
    block = (bz * counts[1] + by) * counts[0] + bx
    order = np.argsort(block, kind="stable")
    block, sid = block[order], sid[order]
    bounds = np.searchsorted(block, np.arange(counts.prod() + 1))
    for bi in range(int(counts.prod())):
        ids = sid[bounds[bi]:bounds[bi + 1]]
        if len(ids) == 0:
            continue
        b = np.array([bi % counts[0], (bi // counts[0]) % counts[1], bi // (counts[0] * counts[1])])
        lo = np.array([edges[a][b[a]] for a in range(3)])
        hi = np.array([edges[a][b[a] + 1] for a in range(3)])
        recurse(lo, hi, ids)
    return bins


# ---- synthetic inputs ---------------------------------------------------

def morton_order(pos, bits=10):
    """Indices that sort points along a Morton curve of their quantised position (the manual's
    coherence advice, doc/mlsgpu-user-manual.xml:211-216)."""
    p = pos - pos.min(axis=0)
    p = (p / (p.max() + 1e-30) * ((1 << bits) - 1)).astype(np.uint64)
    code = np.zeros(len(pos), np.uint64)
    for b in range(bits):
        for a in range(3):
            code |= ((p[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
    return np.argsort(code, kind="stable")


def sphere_splats(n, seed=1, radius=0.005, noise=0.002, R=1.0):
    """C1: noisy unit sphere, uniform area sampling, exact radial normals, constant radius."""
    rng = np.random.Generator(np.random.Philox(seed))
    z = rng.uniform(-1.0, 1.0, n)
    t = rng.uniform(-np.pi, np.pi, n)
    xy = np.sqrt(1.0 - z * z)
    nrm = np.stack([np.cos(t) * xy, np.sin(t) * xy, z], axis=1)
    rad = R + rng.normal(0.0, noise, n)
    out = np.zeros(n, SPLAT_DTYPE)
    out["position"] = (nrm * rad[:, None]).astype(np.float32)
    out["normal"] = nrm.astype(np.float32)
    out["radius"] = np.float32(radius)
    out = out[morton_order(out["position"])]
    return np.ascontiguousarray(out)


def torus_splats(n, seed=2, R=1.0, r=0.4, bump=0.02, freq=20.0):
    """C2: bumpy torus with analytic normals and per-splat radii = local spacing x U(1,2)."""
    rng = np.random.Generator(np.random.Philox(seed))
    # area-uniform sampling of the torus: density in v proportional to (R + r cos v)
    u = rng.uniform(0.0, 2 * np.pi, n)
    v = rng.uniform(0.0, 2 * np.pi, int(n * 1.6) + 16)
    keep = rng.uniform(0.0, 1.0, len(v)) < (R + r * np.cos(v)) / (R + r)
    v = v[keep][:n]
    if len(v) < n:
        v = np.concatenate([v, rng.uniform(0.0, 2 * np.pi, n - len(v))])
    rr = r + bump * np.sin(freq * u) * np.sin(freq * v)

    def point(u, v, rr):
        return np.stack([(R + rr * np.cos(v)) * np.cos(u), (R + rr * np.cos(v)) * np.sin(u), rr * np.sin(v)], axis=1)

    p = point(u, v, rr)
    # analytic normal = dP/du x dP/dv
    drr_du = bump * freq * np.cos(freq * u) * np.sin(freq * v)
    drr_dv = bump * freq * np.sin(freq * u) * np.cos(freq * v)
    cu, su, cv, sv = np.cos(u), np.sin(u), np.cos(v), np.sin(v)
    dpu = np.stack([-(R + rr * cv) * su + drr_du * cv * cu, (R + rr * cv) * cu + drr_du * cv * su, drr_du * sv], axis=1)
    dpv = np.stack([(drr_dv * cv - rr * sv) * cu, (drr_dv * cv - rr * sv) * su, drr_dv * sv + rr * cv], axis=1)
    nrm = np.cross(dpu, dpv)
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    area = 4 * np.pi * np.pi * R * r
    spacing = np.sqrt(area / n)
    out = np.zeros(n, SPLAT_DTYPE)
    out["position"] = p.astype(np.float32)
    out["normal"] = nrm.astype(np.float32)
    out["radius"] = (spacing * rng.uniform(1.0, 2.0, n)).astype(np.float32)
    out = out[morton_order(out["position"])]
    return np.ascontiguousarray(out)


def _torus_points(rng, n, u_lo, u_hi, R, r, bump, freq, spacing):
    """n area-uniform samples of the bumpy torus with u in [u_lo, u_hi): positions, unit normals, radii."""
    u = rng.uniform(u_lo, u_hi, n)
    v = rng.uniform(0.0, 2 * np.pi, int(n * 1.6) + 16)
    keep = rng.uniform(0.0, 1.0, len(v)) < (R + r * np.cos(v)) / (R + r)
    v = v[keep][:n]
    if len(v) < n:
        v = np.concatenate([v, rng.uniform(0.0, 2 * np.pi, n - len(v))])
    rr = r + bump * np.sin(freq * u) * np.sin(freq * v)
    cu, su, cv, sv = np.cos(u), np.sin(u), np.cos(v), np.sin(v)
    p = np.stack([(R + rr * cv) * cu, (R + rr * cv) * su, rr * sv], axis=1)
    drr_du = bump * freq * np.cos(freq * u) * np.sin(freq * v)
    drr_dv = bump * freq * np.sin(freq * u) * np.cos(freq * v)
    dpu = np.stack([-(R + rr * cv) * su + drr_du * cv * cu, (R + rr * cv) * cu + drr_du * cv * su, drr_du * sv], axis=1)
    dpv = np.stack([(drr_dv * cv - rr * sv) * cu, (drr_dv * cv - rr * sv) * su, drr_dv * sv + rr * cv], axis=1)
    nrm = np.cross(dpu, dpv)
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    out = np.zeros(n, SPLAT_DTYPE)
    out["position"] = p.astype(np.float32)
    out["normal"] = nrm.astype(np.float32)
    out["radius"] = (spacing * rng.uniform(1.0, 2.0, n)).astype(np.float32)
    return out


class ShardedTorus:
    """ONE bumpy-torus scan of `n` splats (the C2 surface) laid out so that each of `world`
    GPUs can build exactly its share without generating the rest: the scan consists of `chunks`
    angular sectors of n / chunks samples each (sector c = u in [2 pi c / chunks, 2 pi (c+1) / chunks), its own
    Philox stream, Morton-ordered inside), the full grid is cut into the same blocks of <= max_cells
    cells on every rank (make_bins' first cut), the blocks are dealt to the ranks in contiguous angular
    runs of equal estimated load (owners()), and a rank generates only the sectors its blocks can reach.  The bins a rank ends up
    with are identical to the bins of those blocks had the whole scan been binned in one place
    (tests/test_multi_gpu_host.py checks that), so N ranks mesh ONE surface with no collective and
    no replicated work — the static spatial partition SURVEY 8e allows for reproducible runs."""

    def __init__(self, n, world, fit_grid, seed=2, chunks=None, R=1.0, r=0.4, bump=0.02, freq=20.0, smooth=4.0,
                 max_cells=255, max_splats=1 << 21, bucket_size=63):
        self.n, self.world, self.seed = int(n), int(world), seed
        self.R, self.r, self.bump, self.freq, self.smooth = R, r, bump, freq, smooth
        self.chunks = int(chunks or max(8 * world, 16))
        self.per_chunk = self.n // self.chunks
        self.n = self.per_chunk * self.chunks
        self.spacing = np.float32(fit_grid)
        self.max_cells, self.max_splats = max_cells, max_splats
        self.sample_spacing = np.sqrt(4 * np.pi * np.pi * R * r / self.n)
        # the full grid, from the surface's analytic bounds (every rank must agree without seeing all splats)
        rmax = float(np.float32(2.0 * self.sample_spacing) * np.float32(smooth))
        bound = np.array([R + r + bump + rmax, R + r + bump + rmax, r + bump + rmax])
        lo = np.floor(-bound / fit_grid).astype(np.int64)
        self.lo = (lo // bucket_size) * bucket_size
        self.hi = np.ceil(bound / fit_grid).astype(np.int64)
        self.num_cells = self.hi - self.lo
        self.rmax_cells = rmax / fit_grid
        counts = -(-self.num_cells // max_cells)
        self.block_counts = counts
        self.edges = [np.arange(counts[a] + 1) * self.num_cells[a] // counts[a] for a in range(3)]
        self._owners = None

    def chunk(self, c):
        """Sector c of the scan, world coordinates (unconditioned)."""
        rng = np.random.Generator(np.random.Philox(key=self.seed, counter=[c, 0, 0, 0]))
        s = _torus_points(rng, self.per_chunk, 2 * np.pi * c / self.chunks, 2 * np.pi * (c + 1) / self.chunks,
                          self.R, self.r, self.bump, self.freq, self.sample_spacing)
        return np.ascontiguousarray(s[morton_order(s["position"])])

    def block_of(self, offset):
        return tuple(int(np.searchsorted(self.edges[a], offset[a], side="right") - 1) for a in range(3))

    def owners(self):
        """{block (bx, by, bz): rank}.  Every rank computes the same map from a coarse sample of the
        whole scan (its own Philox stream, ~300 k points, a fraction of a second): the blocks are
        ordered by the angle of their centre around the torus' axis and the ordered list is cut into
        `world` contiguous runs of equal estimated load — contiguous, so that a rank needs only the
        sectors around its run; equal, because the step time of a sharded run is its slowest rank's."""
        if self._owners is not None:
            return self._owners
        rng = np.random.Generator(np.random.Philox(key=self.seed, counter=[1 << 40, 0, 0, 0]))
        m = min(self.n, 40_000 * max(self.world, 8))
        s = _torus_points(rng, m, 0.0, 2 * np.pi, self.R, self.r, self.bump, self.freq, self.sample_spacing)
        g = world_to_grid(condition_splats(s, self.smooth), self.spacing, self.lo)
        load = {}
        for b in make_bins(g, self.num_cells, self.max_cells, 1 << 30):
            k = self.block_of(b["offset"])
            load[k] = load.get(k, 0) + len(b["ids"])

        def angle(k):
            c = [0.5 * (self.edges[a][k[a]] + self.edges[a][k[a] + 1]) + self.lo[a] for a in range(2)]
            return float(np.arctan2(c[1], c[0]) % (2 * np.pi))

        # every block gets an owner (a block the sample missed may still hold a few real splats)
        blocks = [(bx, by, bz) for bz in range(int(self.block_counts[2])) for by in range(int(self.block_counts[1]))
                  for bx in range(int(self.block_counts[0]))]
        blocks.sort(key=lambda k: (angle(k), k))
        total = float(sum(load.values()))
        self._owners, acc = {}, 0.0
        for k in blocks:
            w = load.get(k, 0)
            self._owners[k] = min(int((acc + 0.5 * w) / total * self.world), self.world - 1)
            acc += w
        self._loads = load
        return self._owners

    def block_chunks(self, bx, by):
        """Sectors that can hold a splat reaching block column (bx, by): the column's footprint, grown by
        the largest radius, is cut into 8 x 8 pieces; a piece counts if it meets the annulus the torus
        projects to, and then contributes the angles its corners span (exact for a rectangle that misses
        the axis; a piece around the axis lies inside the annulus' hole and is skipped)."""
        m = self.rmax_cells + 1.0
        X0, X1 = self.edges[0][bx] + self.lo[0] - m, self.edges[0][bx + 1] + self.lo[0] + m
        Y0, Y1 = self.edges[1][by] + self.lo[1] - m, self.edges[1][by + 1] + self.lo[1] + m
        rho_min = (self.R - self.r - self.bump) / float(self.spacing)
        rho_max = (self.R + self.r + self.bump) / float(self.spacing)
        w = 2 * np.pi / self.chunks
        need = set()
        for i in range(8):
            for j in range(8):
                x0, x1 = X0 + (X1 - X0) * i / 8, X0 + (X1 - X0) * (i + 1) / 8
                y0, y1 = Y0 + (Y1 - Y0) * j / 8, Y0 + (Y1 - Y0) * (j + 1) / 8
                near = np.hypot(max(x0, -x1, 0.0), max(y0, -y1, 0.0))
                far = np.hypot(max(abs(x0), abs(x1)), max(abs(y0), abs(y1)))
                if far < rho_min or near > rho_max:
                    continue
                if x0 <= 0.0 <= x1 and y0 <= 0.0 <= y1:
                    return set(range(self.chunks))        # cannot happen for a torus with a hole wider than a piece
                mid = np.arctan2(0.5 * (y0 + y1), 0.5 * (x0 + x1))
                d = [((np.arctan2(y, x) - mid + np.pi) % (2 * np.pi)) - np.pi for x, y in ((x0, y0), (x0, y1), (x1, y0), (x1, y1))]
                c0, c1 = int(np.floor((mid + min(d)) / w)), int(np.floor((mid + max(d)) / w))
                need |= {c % self.chunks for c in range(c0, c1 + 1)}
        return need

    def rank_share(self, rank):
        """(grid splats, bins, lo, spacing) of `rank`: only its blocks' bins, complete."""
        mine = {k for k, r in self.owners().items() if r == rank}
        need = set()
        for bx, by in {(k[0], k[1]) for k in mine}:
            need |= self.block_chunks(bx, by)
        parts = [self.chunk(c) for c in sorted(need)]
        s = np.concatenate(parts) if parts else np.zeros(0, SPLAT_DTYPE)
        s = condition_splats(s, self.smooth)
        g = world_to_grid(s, self.spacing, self.lo)
        bins = make_bins(g, self.num_cells, self.max_cells, self.max_splats)
        return g, [b for b in bins if self.block_of(b["offset"]) in mine], self.lo, self.spacing


def prepare(splats, fit_grid, smooth=4.0, max_cells=255, max_splats=1 << 21, bucket_size=63):
    """World-space splats -> (grid_splats, bins, lo, spacing): everything the reference does on
    the host before DeviceWorkerGroup sees a bin."""
    s = condition_splats(splats.copy(), smooth)
    lo, hi = bounding_grid(s, fit_grid, bucket_size)
    g = world_to_grid(s, fit_grid, lo)
    bins = make_bins(g, hi - lo, max_cells, max_splats)
    return g, bins, lo, np.float32(fit_grid)


def parse_statistics(text):
    """{name: value} from a `mlsgpu --statistics` listing (the reference's Statistics::Registry stream,
    src/statistics.cpp:99-101,175-182,206-208): counters and peaks are ints; variables
    ("sum : mean [+/- stddev] [n]") become {"sum", "mean", "stddev", "n"}; header lines stay strings."""
    out = {}
    for line in text.splitlines():
        name, sep, value = line.partition(": ")
        if not sep:
            continue
        value = value.strip()
        if value.endswith("]") and "[" in value:
            body, _, n = value[:-1].rpartition("[")
            var = {"n": int(n), "sum": 0.0, "mean": None, "stddev": None}
            if body.strip():
                head, _, dev = body.partition("+/-")
                total, _, mean = head.partition(" : ")
                var["sum"], var["mean"] = float(total), float(mean)
                if dev.strip():
                    var["stddev"] = float(dev)
            out[name] = var
        else:
            try:
                out[name] = int(value)
            except ValueError:
                out[name] = value
    return out


def partition_bins(bins, world):
    """Static partition of bins over `world` GPUs (one worker group per GPU, no collective on the
    data path): longest-processing-time greedy on the splat count, deterministic.  The
    reference hands bins to the device with the most free staging capacity
    (src/workers.cpp:329-343); a static split makes scaling runs reproducible.
    Returns a list of `world` lists of bin indices."""
    order = sorted(range(len(bins)), key=lambda i: (-len(bins[i]["ids"]), i))
    loads = [0] * world
    parts = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += len(bins[i]["ids"])
    return [sorted(p) for p in parts]


def bind_to_gpu_numa_node(device):
    """Pins the calling process to the CPUs of the NUMA node the GPU hangs off, so that worker
    threads and the pinned staging / mesh buffers they first touch are local to the GPU's PCIe root.
    With one process per GPU on a multi-socket box this keeps the mesh read-back (about twice the
    input in bytes) off the inter-socket link.  Returns the node, or None if it cannot be found."""
    import os
    import subprocess
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(device), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return None
        if len(bus.split(":")[0]) == 8:       # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, subprocess.SubprocessError):
        return None
