"""Multi-GPU host logic on CPU: bins are independent units, sharded statically over ranks with no
collective on the data path; ranks only meet in the timing all-reduce of bench.py.  Run with the
gloo backend and world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mlsgpu_b200 import host


def make_bins():
    s = host.sphere_splats(30000, seed=6, radius=0.02)
    g, bins, lo, spacing = host.prepare(s, 0.02, max_cells=63, max_splats=4000)
    return g, bins


def test_partition_is_disjoint_cover_and_balanced():
    g, bins = make_bins()
    for world in (1, 2, 4, 8):
        parts = host.partition_bins(bins, world)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(bins)))
        loads = [sum(len(bins[i]["ids"]) for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(len(b["ids"]) for b in bins)
    assert host.partition_bins(bins, 2) == host.partition_bins(bins, 2)   # deterministic


def _rank_main(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, bins = make_bins()
    mine = host.partition_bins(bins, world)[rank]
    # what bench.py reduces: units processed (SUM) and elapsed time (MAX over ranks)
    units = torch.tensor([sum(len(bins[i]["ids"]) for i in mine), len(mine)], dtype=torch.float64)
    elapsed = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(units, op=dist.ReduceOp.SUM)
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        total = sum(len(b["ids"]) for b in bins)
        out.put((units.tolist(), elapsed.item(), total, len(bins)))
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    units, elapsed, total, nbins = out.get()
    assert units == [float(total), float(nbins)]
    assert elapsed == 2.0


# ---- the sharded workload bench.py runs on N > 1 GPUs: ONE scan, every rank builds only its share ----

def _bin_table(g, bins):
    return {(tuple(int(v) for v in b["offset"]), tuple(int(v) for v in b["size"])): g[b["ids"]].tobytes() for b in bins}


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_torus_shares_are_the_bins_of_the_whole_scan(world):
    """The ranks' shares are disjoint, complete and identical, splat for splat and in order, to binning
    the whole scan in one place; the estimated-load partition balances the ranks within a few percent;
    a rank generates only the sectors its blocks can reach."""
    st = host.ShardedTorus(400_000, world, 0.0016, chunks=8 * world)
    whole = np.concatenate([st.chunk(c) for c in range(st.chunks)])
    g = host.world_to_grid(host.condition_splats(whole, st.smooth), st.spacing, st.lo)
    want = _bin_table(g, host.make_bins(g, st.num_cells, st.max_cells, st.max_splats))
    got, loads, generated = {}, [], []
    for rank in range(world):
        gr, bins, lo, spacing = st.rank_share(rank)
        assert np.array_equal(lo, st.lo) and spacing == st.spacing
        share = _bin_table(gr, bins)
        assert not (set(share) & set(got))
        got.update(share)
        loads.append(sum(len(b["ids"]) for b in bins))
        generated.append(len(gr))
    assert got == want
    assert max(loads) / np.mean(loads) < 1.08
    assert sum(generated) < world * st.n        # the ranks did not all generate the whole scan


def _sharded_rank(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = host.ShardedTorus(200_000, world, 0.005, chunks=8 * world)
    g, bins, lo, spacing = st.rank_share(rank)
    # what bench.py reduces for the sharded run: bytes moved (SUM), elapsed (MAX); the splat total is the scan's
    t = torch.tensor([float(sum(len(b["ids"]) for b in bins)), float(len(bins))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        out.put((t.tolist(), st.n))
    dist.destroy_process_group()


def test_sharded_torus_two_ranks_gloo():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_sharded_rank, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    (instances, nbins), n = out.get()
    st = host.ShardedTorus(200_000, world, 0.005, chunks=16)
    whole = np.concatenate([st.chunk(c) for c in range(st.chunks)])
    g = host.world_to_grid(host.condition_splats(whole, st.smooth), st.spacing, st.lo)
    full = host.make_bins(g, st.num_cells, st.max_cells, st.max_splats)
    assert n == 200_000 and nbins == len(full) and instances == sum(len(b["ids"]) for b in full)
