#!/usr/bin/env python
"""Device-to-host bandwidth of one GPU alone and of all ranks' GPUs at once (run under torchrun).

The per-bin path ends in a mesh readback (HostKeyMesh, ~72 bytes per input splat on the C2 torus);
this probe measures what the platform gives that readback when N processes copy concurrently.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/probe_d2h.py
"""
import os
import time

import torch
import torch.distributed as dist


def measure(dev_buf, host_buf, reps):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        host_buf.copy_(dev_buf, non_blocking=True)
    torch.cuda.synchronize()
    return dev_buf.numel() * reps / (time.perf_counter() - t) / 1e9


def main():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    results = {}
    for mb in (8, 64):
        dev = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
        host = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True)
        hsrc = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True)
        measure(dev, host, 5)
        reps = 2048 // mb
        # every rank alone, one after the other
        alone = 0.0
        for r in range(world):
            if world > 1:
                dist.barrier()
            if r == rank:
                alone = measure(dev, host, reps)
        if world > 1:
            dist.barrier()
        together = measure(dev, host, reps)
        if world > 1:
            dist.barrier()
        # H2D at the same time on another stream, as the pipeline has it
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for _ in range(reps):
                dev2 = dev.new_empty(dev.shape)
                dev2.copy_(hsrc, non_blocking=True)
        duplex = measure(dev, host, reps)
        side.synchronize()
        results[mb] = (alone, together, duplex)
    rows = [results]
    if world > 1:
        rows = [None] * world
        dist.all_gather_object(rows, results)
    if rank == 0:
        for mb in (8, 64):
            a = [r[mb][0] for r in rows]
            t = [r[mb][1] for r in rows]
            d = [r[mb][2] for r in rows]
            print("D2H %2d MB copies, %d GPU(s): alone %.1f GB/s per GPU; all at once %.1f GB/s per GPU (%.1f total); "
                  "all at once with H2D running %.1f per GPU" % (mb, world, sum(a) / world, sum(t) / world, sum(t), sum(d) / world))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
