"""Stage-by-stage parity driver for the CUDA path (through the C ABI) against a recorded reference step.

A "record" is what pinlib.record_step() produces from the compiled reference (live, or loaded from tests/golden/):
the state after every phase of one time_step.  Each device stage is fed the recorded INPUT of that stage and its
output is compared with the recorded output, so errors do not chain and integer quantities can be held exact.
"""
import numpy as np

import pinlib as PL
from libfluid_b200 import capi

# tolerances (fp64 device arithmetic, same inputs): see DESIGN.md "Parity"
TOL_P2G = 1e-12        # rel-L2 of post-P2G face velocities
TOL_RHS = 1e-12        # rel-L2 of b
TOL_PRESSURE = 1e-5    # rel-L2 of p where |b| >> tolerance (both solvers stop at an ABSOLUTE |r| < 1e-6)
TOL_FACES = 1e-12      # rel-L2 of faces after apply_pressure / extrapolation given the same pressure
TOL_PARTICLE = 1e-12   # rel-L2 of positions / velocities / c given the same inputs


def context_for(meta_or_ref, **kw):
    if isinstance(meta_or_ref, dict):
        m = meta_or_ref
        size, h, off, g = m["meta/size"], float(m["meta/h"]), m["meta/offset"], m["meta/gravity"]
        method, blend = int(m["meta/method"]), float(m["meta/blend"])
    else:
        r = meta_or_ref
        size, h, off, g, method, blend = r.size, r.h, r.offset, r.gravity, r.method, r.blend
    params = dict(cell_size=h, grid_offset=off, gravity=g, method=method, blending_factor=blend)
    params.update(kw)
    return capi.Context([int(s) for s in size], **params)


def canon(parts):
    """canonical particle order: by cell key, then position, then velocity and c rows (the reference's in-cell order is
    unspecified; particles clamped into the same wall corner coincide and differ only in their payload)"""
    p, v = parts["position"], parts["velocity"]
    keys = [parts[f][:, k] for f in ("cz", "cy", "cx") for k in (2, 1, 0)]
    keys += [v[:, 2], v[:, 1], v[:, 0], p[:, 2], p[:, 1], p[:, 0], parts["raw_cell_index"]]
    return np.lexsort(tuple(keys))


def check_device_against_record(ctx, rec, orc, exact=False):
    """Returns {stage: error} for every stage that is out of tolerance (empty dict == parity)."""
    dt = float(rec["dt"])
    flip = int(ctx.params.method) == capi.FLIP
    bad = {}

    def need(name, ok, detail=""):
        if not ok:
            bad[name] = detail

    def close(name, a, b, tol):
        if exact:
            need(name, np.array_equal(a, b), "not bit-exact: %g" % PL.rel_l2(a, b))
        else:
            e = PL.rel_l2(a, b)
            need(name, e <= tol, "rel_l2 %.3e > %.1e" % (e, tol))

    # ---- A1 advect ----
    ctx.upload_cells(rec["hash0/cells"])
    ctx.upload_particles(rec["hash0/particles"])
    ctx.advect(dt)
    out = ctx.download_particles()
    close("advect", out["position"], rec["advect/particles"]["position"], 0.0 if exact else 1e-15)
    need("advect_old", np.array_equal(out["old_position"], rec["advect/particles"]["old_position"]))
    # ---- A2 collide #1 ----
    ctx.upload_particles(rec["advect/particles"])
    ctx.collide()
    out = ctx.download_particles()
    close("collide1", out["position"], rec["collide1/particles"]["position"], TOL_PARTICLE)
    need("collide1_old", np.array_equal(out["old_position"], out["position"]))
    # ---- K1/K2 keys, sort, table ----
    ctx.upload_particles(rec["collide1/particles"])
    ctx.hash()
    out = ctx.download_particles()
    refh = rec["hash/particles"]
    need("keys", np.array_equal(out["raw_cell_index"], refh["raw_cell_index"]), "sorted key sequence differs")
    a, b = out[canon(out)], refh[canon(refh)]
    need("sort_payload", all(np.array_equal(a[f], b[f]) for f in ("position", "velocity", "cx", "cy", "cz")),
         "particle multiset per cell differs")
    begin, count = ctx.download_table()
    need("table", np.array_equal(count, rec["hash/count"]) and np.array_equal(begin, rec["hash/begin"]))
    need("fluid_cells", np.array_equal(ctx.download_fluid_cells(), rec["hash/fluid_cells"]))
    # ---- P1-P3 P2G (fed in the reference's own sorted order; the device sort is stable so it is kept) ----
    ctx.upload_cells(rec["hash/cells"])
    ctx.upload_particles(refh)
    ctx.hash()
    need("stable_sort", np.array_equal(ctx.download_particles()["position"], refh["position"]))
    ctx.p2g()
    cells = ctx.download_cells()
    need("classification", np.array_equal(cells["type"], rec["p2g/cells"]["type"]))
    close("p2g", cells["vel"], rec["p2g/cells"]["vel"], TOL_P2G)
    if flip:
        close("p2g_old", ctx.download_old_cells()["vel"], rec["p2g/old_cells"]["vel"], TOL_P2G)
    # ---- G0 gravity ----
    ctx.upload_cells(rec["p2g/cells"])
    ctx.gravity(dt)
    close("gravity", ctx.download_cells()["vel"], rec["gravity/cells"]["vel"], 1e-15)
    # ---- S1-S3 system ----
    ctx.upload_cells(rec["gravity/cells"])
    bvec, flags = ctx.download_rhs(dt)
    need("flags", np.array_equal(flags, rec["rhs/flags"]))
    close("rhs", bvec, rec["rhs/b"], TOL_RHS)
    # ---- S6 A*v, against the restatement's _apply_a on a random vector ----
    fluid = rec["hash/fluid_cells"]
    gv, ty = PL.split_cells(rec["gravity/cells"])
    imap, oflags, ob = orc.solver_setup(gv, ty, fluid)
    a_scale = dt / (orc.P.rho * orc.P.h * orc.P.h)
    rng = np.random.default_rng(7)
    v = rng.standard_normal(fluid.shape[0])
    close("apply_a", ctx.apply_a(dt, v), orc.apply_a(a_scale, fluid, imap, oflags, v), 1e-14)
    # ---- S4-S8 solve: residual in the reference's scaling + pressure against the reference's ----
    res, iters = ctx.pressure_solve(dt)
    p = ctx.download_pressure()
    if int(rec["solve/iters"]) == 0:
        need("solve_early_out", iters == 0 and res == 0.0 and not p.any(), "iters %d res %g" % (iters, res))
    else:
        A = lambda x: orc.apply_a(a_scale, fluid, imap, oflags, x)  # noqa: E731
        true_res = np.abs(ob - A(p)).max()
        need("solve_residual", res < ctx.params.tolerance and true_res < 2e-6,
             "reported %.3e true %.3e iters %d" % (res, true_res, iters))
        # Both solvers stop on an ABSOLUTE residual of 1e-6, so the two pressures agree to within their stopping
        # criteria: |A (p_dev - p_ref)|_inf <= |r_dev|_inf + |r_ref|_inf.  Where the system is well scaled
        # (|b| >> tolerance) that also pins p itself to rel-L2 1e-6.
        pref = rec["solve/p"]
        ref_res = np.abs(ob - A(pref)).max()
        gap = np.abs(A(p - pref)).max()
        need("pressure_residual_gap", gap <= true_res + ref_res + 1e-9, "%.3e > %.3e + %.3e" % (gap, true_res, ref_res))
        if np.abs(ob).max() > 1.0:
            close("pressure", p, pref, TOL_PRESSURE)
        # post-projection faces with the device's own pressure against the reference's (1e-7)
        ctx.apply_pressure(dt)
        e = PL.rel_l2(ctx.download_cells()["vel"], rec["apply_pressure/cells"]["vel"])
        need("faces_after_own_solve", e <= 1e-7, "rel_l2 %.3e (|b|max %.2e)" % (e, np.abs(ob).max()))
        ctx.upload_cells(rec["gravity/cells"])
        bvec, flags = ctx.download_rhs(dt)
    # ---- S9 apply_pressure with the reference's pressure ----
    ctx.upload_pressure(rec["solve/p"])
    ctx.apply_pressure(dt)
    close("apply_pressure", ctx.download_cells()["vel"], rec["apply_pressure/cells"]["vel"], TOL_FACES)
    # ---- A3 correct (same table, same in-cell order) ----
    ctx.upload_particles(refh)
    ctx.hash()
    ctx.correct(dt)
    out = ctx.download_particles()
    keep = ~PL.degenerate_mask(np.ascontiguousarray(refh["position"]))
    close("correct", out["position"][keep], rec["correct/particles"]["position"][keep], TOL_PARTICLE)
    # ---- A2 collide #2 ----
    ctx.upload_particles(rec["correct/particles"])
    ctx.collide()
    close("collide2", ctx.download_particles()["position"], rec["collide2/particles"]["position"], TOL_PARTICLE)
    # ---- E1 extrapolate (needs the particle counts of the sorted table) ----
    ctx.upload_particles(refh)
    ctx.hash()
    ctx.upload_cells(rec["apply_pressure/cells"])
    ctx.extrapolate()
    close("extrapolate", ctx.download_cells()["vel"], rec["extrapolate/cells"]["vel"], TOL_FACES)
    # ---- G1-G4 G2P ----
    ctx.upload_cells(rec["extrapolate/cells"])
    if flip:
        ctx.upload_old_cells(rec["p2g/old_cells"])
    ctx.upload_particles(rec["collide2/particles"])
    ctx.g2p()
    out = ctx.download_particles()
    refg = rec["g2p/particles"]
    close("g2p_velocity", out["velocity"], refg["velocity"], TOL_PARTICLE)
    # c rows are velocity GRADIENTS: where the grid velocity is (nearly) uniform the reference's corner-by-corner sum
    # leaves cancellation noise of ~1e-16 |v| / h, so the error is measured against max(|c|, |v| / h)
    vscale = np.linalg.norm(refg["velocity"]) / float(ctx.params.cell_size)
    for f in ("cx", "cy", "cz"):
        err, scale = np.linalg.norm(out[f] - refg[f]), max(np.linalg.norm(refg[f]), vscale)
        need("g2p_" + f, err <= TOL_PARTICLE * scale, "|err| %.3e > %.1e * %.3e" % (err, TOL_PARTICLE, scale))
    cfl = ctx.cfl()
    need("cfl", abs(cfl - float(rec["g2p/cfl"])) <= 1e-14 * abs(cfl), "%r vs %r" % (cfl, float(rec["g2p/cfl"])))
    return bad
