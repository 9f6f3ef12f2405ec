#!/usr/bin/env python
"""Per-call wall times of the positions-streaming loop (bench.py's e2e.positions_streaming leg) at 256^3."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libfluid_b200 import capi  # noqa: E402
import bench as B  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = capi.Context((n, n, n), cell_size=1.0, gravity=B.GRAVITY, method=capi.APIC, max_iterations=1000)
for k, (start, size) in enumerate(B.scene_boxes(n, n, n)):
    ctx.seed_box_device(start, size, density=2, seed=20261017, append=k > 0)
npart = ctx.num_particles()
pin = capi.PinnedBuffer(npart * 24)
for _ in range(3):
    ctx.time_step()
ctx.sync()
for it in range(6):
    t0 = time.perf_counter()
    ctx.time_step()
    t1 = time.perf_counter()
    ctx.wait_transfers()
    t2 = time.perf_counter()
    ctx.download_positions_async(pin.ptr.value, npart)
    t3 = time.perf_counter()
    print("iter %d: step %.1f ms, wait %.1f ms, issue %.1f ms" % (it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)), flush=True)
t0 = time.perf_counter()
ctx.wait_transfers()
print("final wait %.1f ms" % (1e3 * (time.perf_counter() - t0)))
t0 = time.perf_counter()
ctx.download_positions_async(pin.ptr.value, npart)
ctx.wait_transfers()
print("isolated download %.1f ms (%.1f GB/s)" % (1e3 * (time.perf_counter() - t0), npart * 24 / (time.perf_counter() - t0) / 1e9))
