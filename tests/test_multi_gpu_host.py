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
