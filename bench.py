#!/usr/bin/env python
"""bench.py -- particle-substeps/s of a full libfluid time step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            the CUDA path (this repo)
  python bench.py --impl reference --steps K --warmup W     the reference's own CPU code on the host cores

A "step" is one simulation::time_step() (CFL-limited dt, src/simulation.cpp:127-129) over the resident scene:
advect + collide, cell sort, P2G (+gravity), pressure solve, pressure update, position correction + collide,
extrapolation, G2P.  N = 1: 256^3 grid, APIC, 8 particles per cell, ~130 M particles (BASELINE configs[2], the
configuration the metric is quoted on).  N > 1: weak scaling with 256^3 cells per GPU, z-slab decomposition:
512x256x256 on 2 GPUs, 512x512x256 on 4, 512^3 on 8 (BASELINE configs[3]).
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-substeps/s at 256^3 APIC (8 ppc)"
UNIT = "particle-substeps/s"
GRAVITY = (0.0, -981.0, 0.0)


def scene_boxes(nx, ny, nz):
    """tidal step: water fills the box up to y = ny - 2 on the left half and ny - 14 (scaled) on the right half"""
    hi, lo = ny - max(2, ny // 128), ny - max(3, (14 * ny) // 256)
    return [((0.0, 0.0, 0.0), (nx / 2.0, float(hi), float(nz))), ((nx / 2.0, 0.0, 0.0), (nx / 2.0, float(lo), float(nz)))]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.rows, self.proc, self.stopped = gpu_index, [], None, False

    def run(self):
        for q in (self.Q, self.Q.replace("clocks_event_reasons", "clocks_throttle_reasons")):
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + q,
                                              "--format=csv,noheader,nounits", "-lms", "200"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                for line in self.proc.stdout:
                    cols = [c.strip() for c in line.split(",")]
                    if len(cols) >= 8:
                        self.rows.append(cols)
            except Exception:
                pass
            if self.rows or self.stopped:
                break

    def stop(self):
        self.stopped = True
        if self.proc:
            self.proc.terminate()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": len(self.rows)}
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        if sm:
            loaded = sorted(sm)[len(sm) // 4:]  # drop the idle quarter at either end of the window
            out["sm_mhz"] = statistics.median(loaded)
            out["sm_max_mhz"] = float(self.rows[0][2])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for k, nm in enumerate(names):
                if any(len(r) >= 8 and r[4 + k].lower().startswith("active") for r in self.rows):
                    out["reasons"].append(nm)
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


TRAFFIC_KERNEL = {"p2g": ("k_p2g_march<2, 8>", "k_p2g_march<2>"), "g2p": ("k_g2p<2>",),
                  "correct_collide": ("k_correct_tile<1>", "k_correct_tiled3<1>"), "advect_collide": ("k_advect_collide",)}


def captured_traffic(grid):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the single-kernel phases, from the
    committed ncu capture of this same command at 256^3 (profiles/*_traffic_256.json, tools/ncu_traffic_table.py);
    None for other grids or when no capture is committed."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic_%d.json" % grid)))
    if not files:
        return {}, None
    tab = json.load(open(files[-1]))["kernels"]
    out = {}
    for ph, names in TRAFFIC_KERNEL.items():
        for k in names:
            if k in tab:
                out[ph] = tab[k]["dram_bytes_per_launch"]
                break
    return out, os.path.relpath(files[-1], ROOT)


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, n=None, steps=None, warmup=None, quiet=False):
    """The reference's own CPU implementation (oracle/_ref = unmodified libfluid sources) on the host cores, on a
    bounded sample of the workload: the same tidal-step scene on a smaller grid."""
    from oracle import refbind as RB
    kind = "reference"
    n = n or args.ref_grid
    steps = steps if steps is not None else args.steps
    warmup = warmup if warmup is not None else args.warmup
    cores = int(os.environ.get("LFK_REF_THREADS", "0")) or os.cpu_count() or 1
    # (torchrun exports OMP_NUM_THREADS=1 to its workers: the reference's OpenMP phases would silently run on one core)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    if not RB.available():
        raise SystemExit("oracle/_ref/libfluid_ref.so is missing (built by __graft_entry__.build() where "
                         "/root/reference exists)")
    sim = RB.RefSim((n, n, n), gravity=GRAVITY, method=RB.APIC)
    for start, size in scene_boxes(n, n, n):
        sim.seed_box(start, size)
    sim.reset_space_hash()
    npart = sim.num_particles()
    for _ in range(warmup):
        sim.time_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.time_step()
    dt = time.perf_counter() - t0
    value = npart * steps / dt
    sample = "tidal-step scene at %d^3 (%d particles, 8 ppc), %d steps of simulation::time_step()" % (n, npart, steps)
    base = {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
    if quiet:
        return base
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "256^3 APIC tidal-step full time_step; reference CPU arm runs " + sample},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return base


# ------------------------------------------------------------------------------------------------------ our arm
def workload_dims(n, world):
    """Weak scaling towards BASELINE configs[3]: one axis doubles per doubling of the GPU count, so that 8 GPUs run
    512^3 (64-cell z-slabs) and every GPU always owns n^3 cells: 1: n^3, 2: 2n x n x n, 4: 2n x 2n x n, 8: (2n)^3.
    Other rank counts: n x n x (n * world).  z is the slab axis."""
    if world in (1, 2, 4, 8):
        return n * (2 if world >= 2 else 1), n * (2 if world >= 4 else 1), n * (2 if world >= 8 else 1)
    return n, n, n * world


def run_ours(args):
    import faulthandler
    import traceback
    # a hang (a rank waiting in a collective for a peer that died) must end the run, not sit until the driver's limit
    faulthandler.dump_traceback_later(int(os.environ.get("BENCH_WATCHDOG_S", "780")), exit=True)
    try:
        _run_ours(args)
    except BaseException as ex:  # one rank failing alone would leave its peers inside NCCL: die hard, torchrun reaps them
        if isinstance(ex, SystemExit) and ex.code in (0, None):
            raise
        sys.stderr.write("[rank %s] bench.py failed:\n%s\n" % (os.environ.get("RANK", "0"), traceback.format_exc()))
        sys.stderr.flush()
        os._exit(1)


def _run_ours(args):
    import torch
    from libfluid_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        box = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]

    n = args.grid
    nx, ny, nz = workload_dims(n, world)
    # a real (non-default) torch stream: lfk launches on it and torch.cuda.Event times on it
    tstream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    ctx = capi.Context((nx, ny, nz), device=local_rank, stream=stream, nranks=world, rank=rank, nccl_id=nccl_id,
                       cell_size=1.0, gravity=GRAVITY, method=capi.APIC, max_iterations=args.max_iterations,
                       preconditioner=capi.PRECOND_MULTIGRID)
    z0, z1 = ctx.slab()
    for k, (start, size) in enumerate(scene_boxes(nx, ny, nz)):
        ctx.seed_box_device(start, size, density=2, seed=20261017, append=k > 0)
    np_local = ctx.num_particles()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allsum(v):
        if dist is None:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    def allmax(v):
        if dist is None:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def agree(ok, what):
        """every rank learns whether ANY rank failed a local check before the next collective is entered"""
        from libfluid_b200 import slabs
        if not slabs.agree(ok, dist, device="cuda"):
            if not ok:
                sys.stderr.write("[rank %d] %s\n" % (rank, what))
            raise SystemExit(3)

    np_total = int(allsum(np_local))
    for _ in range(args.warmup):
        ctx.time_step()
    # ---- timed region: device resident, CUDA events on the launching stream, max over ranks ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.reset_stats()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 0
    for _ in range(args.steps):
        ctx.time_step()
        iters += ctx.stats()["pcg_iterations"]
    e1.record()
    barrier()
    ms = allmax(e0.elapsed_time(e1))
    clocks = sampler.stop()
    launches = ctx.stats()["kernel_launches"]
    value = np_total * args.steps / (ms * 1e-3)

    # ---- per-phase device times (separate instrumented steps; the event pairs serialise the phases) ----
    ctx.set_timing(True)
    ctx.reset_stats()
    prof_steps = max(1, min(args.steps, 3))
    prof_iters = 0
    for _ in range(prof_steps):
        ctx.time_step()
        prof_iters += ctx.stats()["pcg_iterations"]
    st = ctx.stats()
    ctx.set_timing(False)
    phase = {k: v / prof_steps for k, v in st["phase_ms"].items() if v > 0}
    # PCG iterations of one step started from p = 0 like the reference (the fused step warm-starts from the last p)
    ctx.set_tuning("warm_start", 0)
    ctx.time_step()
    cold_iters = ctx.stats()["pcg_iterations"]
    ctx.set_tuning("warm_start", 1)
    ctx.time_step()
    np_now = ctx.num_particles()
    nf = ctx.num_fluid_cells()
    ncl = nx * ny * (z1 - z0)
    peak, peak_src = measured_peak()
    # algorithmic bytes per launch (DESIGN.md "Kernels"): single-kernel phases only
    alg = {"p2g": 120 * np_now + 26 * ncl, "g2p": 120 * np_now + 24 * ncl,
           "correct_collide": 48 * np_now + ncl, "advect_collide": 72 * np_now + ncl}
    single = {k: phase[k] for k in alg if k in phase}
    dom = max(single, key=single.get)
    ach = alg[dom] / (single[dom] * 1e-3) / 1e9
    traffic, traffic_src = captured_traffic(n) if world == 1 else ({}, None)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic.get(dom), "traffic_source": traffic_src, "algorithmic_bytes": alg[dom],
                "peak_source": peak_src, "ms": single[dom],
                "all": {k: {"ms": single[k], "GB/s": alg[k] / (single[k] * 1e-3) / 1e9,
                            "frac": alg[k] / (single[k] * 1e-3) / 1e9 / peak, "algorithmic_bytes": alg[k],
                            "traffic": traffic.get(k)} for k in single}}
    if "pcg" in phase and prof_iters:
        it_ms = phase["pcg"] * prof_steps / prof_iters
        roofline["all"]["pcg_iteration"] = {"ms": it_ms, "GB/s": 105 * nf / (it_ms * 1e-3) / 1e9,
                                           "frac": 105 * nf / (it_ms * 1e-3) / 1e9 / peak,
                                           "algorithmic_bytes": 105 * nf,
                                           "iters_per_step": prof_iters / prof_steps, "iters_per_s": 1e3 / it_ms}

    # ---- end to end through the C ABI with HOST buffers: AoS particles + cells up, step, AoS particles + cells down.
    # The particle count of a rank changes from step to step (migration across the slab boundaries), so the count is
    # read every step and the pinned buffer carries headroom; the ranks agree on every local check before the step's
    # collectives are entered.
    e2e = None
    if not args.no_e2e:
        cap = int(np_now * 1.03) + (1 << 16)
        zlo, zhi = max(z0 - 1, 0), min(z1 + 1, nz)
        ok = True
        try:
            host_p = torch.empty(cap * 152, dtype=torch.uint8, pin_memory=True)
            host_c = torch.empty(nx * ny * (zhi - zlo) * 32, dtype=torch.uint8, pin_memory=True)
        except RuntimeError as ex:
            ok, why = False, "pinned host allocation failed: %s" % ex
        agree(ok, why if not ok else "")
        own_off = (z0 - zlo) * nx * ny * 32  # the owned layers inside the slab buffer
        cnt = ctx.download_particles((host_p.data_ptr(), cap))
        ctx.download_cells_slab(host_c.data_ptr() + own_off)
        barrier()
        up_bytes = down_bytes = 0
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            ctx.upload_cells_slab(host_c.data_ptr())
            ctx.upload_particles((host_p.data_ptr(), cnt))
            up_bytes += cnt * 152 + host_c.numel()
            ctx.time_step()
            now = ctx.num_particles()
            agree(now <= cap, "e2e: %d particles exceed the pinned buffer (%d)" % (now, cap))
            cnt = ctx.download_particles((host_p.data_ptr(), cap))
            ctx.download_cells_slab(host_c.data_ptr() + own_off)
            down_bytes += cnt * 152 + ncl * 32
        barrier()
        dt_e2e = allmax(time.perf_counter() - t0)
        e2e = {"value": np_total * args.e2e_steps / dt_e2e, "unit": UNIT,
               "h2d_bytes_per_step": int(up_bytes // args.e2e_steps),
               "d2h_bytes_per_step": int(down_bytes // args.e2e_steps), "steps": args.e2e_steps,
               "bytes_are": "this rank's; every rank moves about as much",
               "what": "lfk_upload_cells_slab + lfk_upload_particles (pinned AoS, reference layouts) + lfk_time_step_cfl"
                       " + lfk_download_particles + lfk_download_cells_slab, per step, wall clock, max over ranks",
               "pcie_floor_s_per_step": (up_bytes + down_bytes) / args.e2e_steps / 55e9}
        del host_p, host_c
        # what a per-frame consumer (renderer, mesher, the Maya node's particle cache) costs: the state stays resident
        # and only the positions stream out, asynchronously on the transfer stream, overlapped with the next step
        pin = None
        try:
            pin = capi.PinnedBuffer(cap * 24)
        except capi.LfkError:
            pin = None
        if allmax(0.0 if pin is not None else 1.0) == 0.0:  # every rank got its pinned buffer
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps + 2):
                ctx.time_step()
                ctx.wait_transfers()
                ctx.download_positions_async(pin.ptr.value, cap)
            ctx.wait_transfers()
            barrier()
            dt_s = allmax(time.perf_counter() - t0)
            e2e["positions_streaming"] = {"value": np_total * (args.e2e_steps + 2) / dt_s, "unit": UNIT,
                                          "d2h_bytes_per_step": int(ctx.num_particles() * 24), "h2d_bytes_per_step": 0,
                                          "what": "lfk_time_step_cfl + lfk_download_positions_async (pinned, 24 B per "
                                                  "particle) per step, the copy overlapped with the next step"}
        else:
            e2e["positions_streaming"] = {"unavailable": "pinned host allocation failed"}
        if pin is not None:
            pin.close()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%dx%dx%d MAC grid (%s; %d^3 cells per GPU, z-slabs of %d layers), APIC, "
                                   "tidal-step scene, 8 ppc, %d particles, full simulation::time_step() with CFL dt"
                                   % (nx, ny, nz, "BASELINE configs[2]" if world == 1 else
                                      ("BASELINE configs[3]" if (nx, ny, nz) == (512, 512, 512) else
                                       "weak-scaling step towards configs[3]"), n, z1 - z0, np_total),
                       "particles": np_total, "fluid_cells_rank0": nf, "pcg_iters_per_step": iters / args.steps,
                       "pcg_iters_cold_start": cold_iters,
                       "l2_policy": "inputs (>= 16 GB of particle state per step) are far larger than the 126 MB L2",
                       "preconditioner": "aggregation multigrid V(2,2), fp32", "tolerance": 1e-6},
            "phase_ms": phase, "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e}
    ctx.close()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = run_reference(args, n=args.ref_grid, steps=2, warmup=1, quiet=True)
                if args.ref_grid_large > args.ref_grid:  # the size trend of the CPU path (one step is ~40 s)
                    line["cpu_baseline"]["larger_sample"] = run_reference(args, n=args.ref_grid_large, steps=1, warmup=1,
                                                                          quiet=True)
            except SystemExit as ex:
                line["cpu_baseline"] = {"unavailable": str(ex)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--ref-grid", type=int, default=64, help="grid of the bounded CPU sample")
    ap.add_argument("--ref-grid-large", type=int, default=128, help="second, larger CPU sample (0: skip)")
    ap.add_argument("--max-iterations", type=int, default=1000)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            run_reference(args)
        return
    run_ours(args)


if __name__ == "__main__":
    main()
