#!/usr/bin/env python
"""A/B timing of kernel variants on the bench scene (runs ON the GPU box).

One process, one resident scene: after a warm-up, every configuration in CONFIGS is applied with lfk_set_tuning and
timed over a few steps with the per-phase CUDA-event timers (lfk_set_timing).  Prints one JSON object per
configuration and writes them all to gpurun_out/<tag>_sweep.json.  Not a bench number: the phase timers serialise
the phases; bench.py is the measurement of record.

  python tools/variant_sweep.py --grid 256 --tag r1c
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NEW = {"p2g": 0, "warm_start": 1, "red_blocks": 0, "graph": 1, "mg_coarse": 0, "lean_sort": 1, "p2g_chunk": 0}  # the library defaults
CONFIGS = [
    ("defaults", dict(NEW)),
    ("defaults+full_sort", dict(NEW, lean_sort=0)),
    ("defaults+no_graph", dict(NEW, graph=0)),
    ("defaults+cold_start", dict(NEW, warm_start=0)),
    ("defaults+p2g_chunk16", dict(NEW, p2g_chunk=16)),
    ("defaults+p2g_chunk32", dict(NEW, p2g_chunk=32)),
    ("defaults+p2g_chunk64", dict(NEW, p2g_chunk=64)),
    ("defaults+p2g_chunk128", dict(NEW, p2g_chunk=128)),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--tag", default="sweep")
    ap.add_argument("--only", default="")
    ap.add_argument("--lib", default="", help="a variant library built by tools/build_variant.py")
    args = ap.parse_args()
    import torch
    from libfluid_b200 import capi
    if args.lib:
        capi.LIB_PATH = os.path.abspath(args.lib)
    import bench as B

    torch.cuda.set_device(0)
    n = args.grid
    ctx = capi.Context((n, n, n), device=0, cell_size=1.0, gravity=B.GRAVITY, method=capi.APIC,
                       max_iterations=1000, preconditioner=capi.PRECOND_MULTIGRID)
    for k, (start, size) in enumerate(B.scene_boxes(n, n, n)):
        ctx.seed_box_device(start, size, density=2, seed=20261017, append=k > 0)
    npart = ctx.num_particles()
    for key, val in NEW.items():
        ctx.set_tuning(key, val)
    for _ in range(args.warmup):
        ctx.time_step()
    out = []
    for name, cfg in CONFIGS:
        if args.only and name not in args.only.split(","):
            continue
        for key, val in cfg.items():
            ctx.set_tuning(key, val)
        try:
            ctx.time_step()  # settle (first warm start, kernel attributes)
            ctx.sync()
            import time
            t0 = time.perf_counter()
            iters = 0
            for _ in range(args.steps):
                ctx.time_step()
                iters += ctx.stats()["pcg_iterations"]
            ctx.sync()
            wall = (time.perf_counter() - t0) * 1e3 / args.steps
            ctx.set_timing(True)
            ctx.reset_stats()
            piters = 0
            for _ in range(args.steps):
                ctx.time_step()
                piters += ctx.stats()["pcg_iterations"]
            st = ctx.stats()
            ctx.set_timing(False)
            phase = {k: round(v / args.steps, 3) for k, v in st["phase_ms"].items() if v > 0}
            row = {"config": name, "ms_per_step_wall": round(wall, 3), "pcg_iters_per_step": iters / args.steps,
                   "phase_ms": phase, "phase_sum": round(sum(phase.values()), 3),
                   "pcg_ms_per_iter": round(phase.get("pcg", 0.0) * args.steps / max(piters, 1), 4),
                   "particles": npart, "residual": st["pcg_residual"]}
        except Exception as ex:  # keep going: the other variants are still worth measuring
            row = {"config": name, "error": repr(ex)}
        print(json.dumps(row), flush=True)
        out.append(row)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", args.tag + "_sweep.json"), "w") as f:
        json.dump(out, f, indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
