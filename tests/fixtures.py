"""Shared fixtures: the reference tests' generators re-expressed in numpy.

Cites are relative to the upstream mlsgpu tree.
"""
import numpy as np

from oracle.oracle import SPLAT_DTYPE


class MT19937:
    """std::tr1::mt19937 with its default seed (5489), 32-bit outputs."""

    def __init__(self, seed=5489):
        mt = [0] * 624
        mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.mt = mt
        self.idx = 624

    def _twist(self):
        mt = self.mt
        for i in range(624):
            y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
            v = mt[(i + 397) % 624] ^ (y >> 1)
            if y & 1:
                v ^= 0x9908B0DF
            mt[i] = v
        self.idx = 0

    def __call__(self):
        if self.idx >= 624:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def uniform_real(self, lo, hi, dtype=np.float64):
        """tr1::variate_generator<mt19937&, uniform_real<T>>: one 32-bit draw scaled to [0,1)."""
        u = dtype(self()) / dtype(4294967296.0)
        return dtype(u * dtype(hi - lo) + dtype(lo))


def sphere_splats(n, center, radius, engine=None):
    """TestMls::sphereSplats, test/test_mls.cpp:349-380."""
    eng = engine or MT19937()
    out = np.zeros(n, SPLAT_DTYPE)
    for i in range(n):
        z = eng.uniform_real(-1.0, 1.0)
        t = eng.uniform_real(-np.pi, np.pi)
        xy = np.sqrt(1.0 - z * z)
        x = np.cos(t) * xy
        y = np.sin(t) * xy
        out[i]["normal"] = (x, y, z)
        out[i]["radius"] = eng.uniform_real(radius, 2.0 * radius)
        out[i]["position"] = (center[0] + x * radius, center[1] + y * radius, center[2] + z * radius)
        out[i]["quality"] = eng.uniform_real(0.0, 1.0)
    return out


def make_splats(rows):
    """rows of (x, y, z, r): addSplat in test/test_splat_tree.cpp:56-70."""
    out = np.zeros(len(rows), SPLAT_DTYPE)
    for i, (x, y, z, r) in enumerate(rows):
        out[i]["position"] = (x, y, z)
        out[i]["radius"] = r
        out[i]["normal"] = (1.0, 0.0, 0.0)
        out[i]["quality"] = 1.0
    return out


def doubles_equal(expected, actual, eps):
    """MLSGPU_ASSERT_DOUBLES_EQUAL, test/testutil.cpp:50-71: NaN-aware, abs-or-rel."""
    if np.isnan(expected) or np.isnan(actual):
        return bool(np.isnan(expected) and np.isnan(actual))
    if expected == actual:
        return True
    err = abs(expected - actual)
    return err <= eps or err <= eps * abs(expected)


def is_manifold(num_vertices, triangles):
    """Manifold::isManifold, test/manifold.h:99-235.  Returns '' or a reason."""
    edges = [[] for _ in range(num_vertices)]
    for t, tri in enumerate(np.asarray(triangles).tolist()):
        idx = list(tri)
        for _ in range(3):
            if idx[0] >= num_vertices:
                return "Triangle %d contains out-of-range index %d" % (t, idx[0])
            if idx[0] == idx[1]:
                return "Triangle %d contains vertex %d twice" % (t, idx[0])
            edges[idx[0]].append((idx[1], idx[2]))
            idx = idx[1:] + idx[:1]
    for i, neigh in enumerate(edges):
        if not neigh:
            return "Vertex %d is isolated" % i
        arrow = {}
        seen = set()
        for x, y in neigh:
            if x in arrow:
                return "Edge %d - %d occurs twice with same winding" % (i, x)
            arrow[x] = y
            if y in seen:
                return "Edge %d - %d occurs twice with same winding" % (y, i)
            seen.add(y)
        length = 0
        for x, _ in neigh:
            if x not in seen:
                cur = x
                while cur in arrow:
                    cur = arrow[cur]
                    length += 1
        if length != 0 and length != len(neigh):
            return "Vertex %d is both in the interior and on the boundary" % i
        if length == 0:
            start = neigh[0][0]
            cur = start
            while True:
                cur = arrow[cur]
                length += 1
                if cur == start:
                    break
            if length != len(neigh):
                return "Vertex %d tunnels between interior regions" % i
    return ""


def walk_chain(start, commands, code, limit=None):
    """Follow the command list from start[code]; returns the list of splat ids visited
    (src/splat_tree.h:47-74 encoding) or raises AssertionError on malformed data."""
    pos = int(start[code])
    out = []
    if pos == -1:
        return out
    limit = limit if limit is not None else len(commands) + 1
    steps = 0
    while True:
        assert 0 <= pos < len(commands)
        end = int(commands[pos])
        pos += 1
        assert pos < end < len(commands)
        out.extend(int(c) for c in commands[pos:end])
        nxt = int(commands[end])
        assert nxt >= -1
        steps += 1
        assert steps <= limit, "infinite loop in command list"
        if nxt == -1:
            return out
        pos = nxt
