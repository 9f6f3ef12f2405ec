#!/usr/bin/env python
"""Two contexts in one process step the same violent scene (the multi-GPU check's); their states must stay bit-identical.
Prints the first step / field at which they differ.  `--one` steps a single context (for compute-sanitizer runs)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libfluid_b200 import capi  # noqa: E402


def make(method, n):
    kw = dict(cell_size=1.0, gravity=(0.0, -981.0, 0.0), method=method, blending_factor=0.95, max_iterations=2000)
    ctx = capi.Context(n, device=0, **kw)
    ctx.seed_box_device((2.0, 1.0, 1.0), (14.0, 9.0, n[2] * 0.45), velocity=(3.0, 0.0, 55.0), density=2, seed=7)
    ctx.seed_box_device((6.0, 3.0, n[2] * 0.55), (15.0, 12.0, n[2] * 0.4), velocity=(-2.0, 0.0, -48.0), density=2,
                        seed=9, append=True)
    return ctx


def main():
    n = (24, 20, 19)
    steps = int(os.environ.get("PROBE_STEPS", "8"))
    if "--one" in sys.argv:
        ctx = make(capi.APIC, n)
        for _ in range(steps):
            ctx.time_step(0.02)
        print("one context:", ctx.num_particles(), "particles,", ctx.stats()["pcg_iterations"], "iterations")
        ctx.close()
        return 0
    bad = 0
    for method in (capi.APIC, capi.FLIP):
        a, b = make(method, n), make(method, n)
        # seeding hands out slots with an atomic counter: make the two contexts start from the SAME array order
        parts = a.download_particles().copy()
        a.upload_particles(parts)
        b.upload_particles(parts)
        for step in range(steps):
            a.time_step(0.02)
            b.time_step(0.02)
            pa, pb = a.download_particles(), b.download_particles()
            ca, cb = a.download_cells(), b.download_cells()
            diffs = [f for f in ("position", "velocity", "cx", "cy", "cz", "raw_cell_index")
                     if not np.array_equal(np.ascontiguousarray(pa[f]).view("u1"), np.ascontiguousarray(pb[f]).view("u1"))]
            if not np.array_equal(np.ascontiguousarray(ca["vel"]).view("u1"), np.ascontiguousarray(cb["vel"]).view("u1")):
                diffs.append("cells.vel")
            if not np.array_equal(ca["type"], cb["type"]):
                diffs.append("cells.type")
            sa, sb = a.stats(), b.stats()
            print("method %d step %d: iters %d / %d, differing: %s" % (method, step, sa["pcg_iterations"],
                                                                      sb["pcg_iterations"], diffs or "none"), flush=True)
            if diffs:
                bad += 1
                for f in diffs:
                    if f.startswith("cells"):
                        continue
                    d = np.abs(pa[f].astype(np.float64) - pb[f].astype(np.float64))
                    d = d.reshape(d.shape[0], -1).max(axis=1)
                    w = np.nonzero(d > 0)[0]
                    print("   %s: %d particles differ, max %.3e, first at index %d pos %s" % (f, w.size, d.max(), w[0],
                                                                                           pa["position"][w[0]]))
                break
        a.close()
        b.close()
    print("determinism probe:", "ok" if bad == 0 else "FAILED")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
