"""Where does worker creation go?  Times context creation, pinned and device allocations and
mlsb_worker_create at two buffer sizes (run on the GPU box)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlsgpu_b200 import capi


def timed(label, fn):
    t = time.time()
    r = fn()
    print("%-44s %.3f s" % (label, time.time() - t), flush=True)
    return r


timed("device_count", capi.device_count)
first = timed("first device alloc (context)", lambda: capi.DeviceBuffer(1 << 20))
for mb in (64, 256):
    b = timed("pinned alloc %d MB" % mb, lambda: capi.PinnedBuffer("u1", mb << 20))
    timed("  free", b.free)
for mb in (64, 512, 2048):
    b = timed("device alloc %d MB" % mb, lambda: capi.DeviceBuffer(mb << 20))
    timed("  free", b.free)
for ms in (150000, 2000000):
    for mesh in (57 << 20, 102 << 20):
        w = timed("worker maxSplats=%d mesh=%d MB" % (ms, mesh >> 20),
                  lambda: capi.Worker(max_splats=ms, max_cells=255, mesh_bytes=mesh))
        timed("  destroy", w.close)
g = timed("group 4 workers maxSplats=2000000 stage", lambda: capi.WorkerGroup(num_workers=4, stage=True, max_splats=2000000, max_cells=255, mesh_bytes=102 << 20))
timed("  destroy", g.close)
g = timed("group 4 workers maxSplats=150000 stage", lambda: capi.WorkerGroup(num_workers=4, stage=True, max_splats=150000, max_cells=255, mesh_bytes=102 << 20))
timed("  destroy", g.close)
