"""GPU parity tests proper: every stage of the CUDA path, called through the C ABI, against
  (a) the golden vectors minted from the unmodified reference (tests/golden/*.npz),
  (b) the compiled reference driven live where oracle/_ref travelled to this box,
  (c) size-independent properties at larger sizes."""
import os

import numpy as np
import pytest

import devlib as DL
import pinlib as PL
from libfluid_b200 import capi
from pinlib import OB, RB

pytestmark = pytest.mark.gpu

def load_golden(scene):
    z = np.load(os.path.join(PL.GOLDEN_DIR, scene + ".npz"))
    rec = {k: z[k] for k in z.files}
    orc = OB.Oracle(rec["meta/size"], h=float(rec["meta/h"]), offset=rec["meta/offset"],
                    gravity=rec["meta/gravity"], method=int(rec["meta/method"]), blend=float(rec["meta/blend"]))
    return orc, rec


@pytest.mark.parametrize("precond", [capi.PRECOND_JACOBI, capi.PRECOND_MULTIGRID])
@pytest.mark.parametrize("scene", PL.SCENES)
def test_stages_match_golden(scene, precond):
    orc, rec = load_golden(scene)
    ctx = DL.context_for(rec, preconditioner=precond, max_iterations=2000)
    assert DL.check_device_against_record(ctx, rec, orc) == {}
    ctx.close()


@pytest.mark.skipif(not RB.available(), reason="oracle/_ref did not travel to this box")
@pytest.mark.parametrize("scene", PL.SCENES)
def test_stages_match_reference_live(scene):
    ref = PL.make_scene(scene, 24)
    orc = PL.oracle_for(ref)
    ctx = DL.context_for(ref, max_iterations=2000)
    for step in range(12):
        dt = min(ref.cfl_number * ref.cfl(), 0.033) if step % 2 else 0.004
        rec = PL.record_step(ref, dt)
        assert DL.check_device_against_record(ctx, rec, orc) == {}, "step %d" % step
    ctx.close()


@pytest.mark.skipif(not RB.available(), reason="oracle/_ref did not travel to this box")
@pytest.mark.parametrize("scene", ["dam_break", "flip_obstacle"])
def test_fused_time_step_tracks_reference(scene):
    """The fused lfk_time_step against the reference's stock time_step from identical state, one step at a time
    (re-synchronised each step: trajectories are chaotic, only the per-step map is comparable)."""
    from scipy.spatial import cKDTree
    ref = PL.make_scene(scene, 20)
    ctx = DL.context_for(ref, max_iterations=2000)
    for step in range(8):
        dt = 0.004 if step < 4 else min(ref.cfl_number * ref.cfl(), 0.033)
        ctx.upload_cells(ref.cells())
        if ref.method == RB.FLIP and step > 0:
            ctx.upload_old_cells(ref.old_cells())
        ctx.upload_particles(ref.particles())
        ref.time_step(dt)
        ctx.time_step(dt)
        a, b = ctx.download_particles(), ref.particles()
        assert a.shape == b.shape
        d, idx = cKDTree(b["position"]).query(a["position"])
        keep = d < 1e-3  # particles on the r^2<1e-12 random-kick branch of the reference are not comparable
        assert keep.mean() > 0.995
        assert np.unique(idx[keep]).size == keep.sum()
        assert np.median(d[keep]) < 1e-6
        assert PL.rel_l2(a["velocity"][keep], b["velocity"][idx[keep]]) < 1e-4
        ca, cb = ctx.download_cells(), ref.cells()
        assert np.array_equal(ca["type"], cb["type"])
    ctx.close()


def test_device_is_run_to_run_deterministic():
    orc, rec = load_golden("dam_break")
    outs = []
    for _ in range(2):
        ctx = DL.context_for(rec, max_iterations=2000)
        ctx.upload_cells(rec["hash0/cells"])
        ctx.upload_particles(rec["hash0/particles"])
        for _ in range(3):
            ctx.time_step(0.004)
        outs.append((ctx.download_particles().copy(), ctx.download_cells().copy()))
        ctx.close()
    assert np.array_equal(outs[0][0].view("u1"), outs[1][0].view("u1"))
    assert np.array_equal(outs[0][1]["vel"], outs[1][1]["vel"])


def test_empty_and_edge_inputs():
    ctx = capi.Context((5, 4, 6), cell_size=0.5, grid_offset=(1, 2, 3), gravity=(0, -9.81, 0))
    ctx.upload_particles(np.zeros(0, dtype=capi.PARTICLE_DTYPE))
    ctx.hash()
    ctx.p2g()
    cells = ctx.download_cells()
    assert (cells["type"] == capi.AIR).all() and not cells["vel"].any()
    res, it = ctx.pressure_solve(0.01)
    assert it == 0 and res == 0.0
    assert ctx.cfl() == np.inf
    ctx.time_step(0.01)
    assert ctx.num_particles() == 0
    # out-of-grid positions clamp into boundary cells (src/simulation.cpp:255-257)
    parts = np.zeros(3, dtype=capi.PARTICLE_DTYPE)
    parts["position"] = [[-5, 2.1, 3.1], [100, 100, 100], [3.49999, 3.99999, 5.99999]]
    ctx.upload_particles(parts)
    ctx.hash()
    keys = np.sort(ctx.download_particles()["raw_cell_index"])
    assert list(keys) == [0, 5 * 4 * 6 - 1, 5 * 4 * 6 - 1]
    fresh = capi.Context((4, 4, 4))
    with pytest.raises(capi.LfkError) as ei:  # lfk_set_params not called yet -> LFK_E_STATE, not a crash
        fresh.hash()
    assert ei.value.code == -2002
    fresh.close()
    ctx.close()


def test_crowded_cell_sort_stays_stable():
    """> 32 particles in one cell takes the block-sort path; order must still be the stable order."""
    n = 8
    ctx = capi.Context((n, n, n), cell_size=1.0)
    rng = np.random.default_rng(3)
    parts = np.zeros(5000, dtype=capi.PARTICLE_DTYPE)
    parts["position"] = rng.uniform(0, n, size=(5000, 3))
    parts["position"][:3000] = rng.uniform(3.0, 4.0, size=(3000, 3))  # 3000 particles in cell (3,3,3)
    parts["velocity"][:, 0] = np.arange(5000)  # identity tag
    ctx.upload_particles(parts)
    ctx.hash()
    out = ctx.download_particles()
    key = OB.Oracle((n, n, n)).cell_keys(np.ascontiguousarray(parts["position"]))
    perm = np.argsort(key, kind="stable")
    assert np.array_equal(out["velocity"][:, 0], parts["velocity"][perm, 0])
    ctx.close()


def test_projection_roundtrip_properties_large():
    """Size-independent properties at 128^3 (config 5): after solve + apply_pressure the discrete divergence of
    every fluid cell is below the solver tolerance; A is symmetric (<Au,v> == <u,Av>) and linear."""
    n = 128
    ctx = capi.Context((n, n, n), cell_size=1.0, max_iterations=2000)
    ctx.synthetic_projection_device(seed=5)
    dt = 1.0 / 60.0
    nf = ctx.num_fluid_cells()
    assert nf == n * (n - 1) * n
    rng = np.random.default_rng(0)
    u, v = rng.standard_normal(nf), rng.standard_normal(nf)
    Au, Av = ctx.apply_a(dt, u), ctx.apply_a(dt, v)
    assert abs(Au @ v - u @ Av) <= 1e-10 * abs(Au @ v)
    assert PL.rel_l2(ctx.apply_a(dt, 2.0 * u - 3.0 * v), 2.0 * Au - 3.0 * Av) < 1e-13
    res, iters = ctx.pressure_solve(dt)
    assert res < 1e-6 and 0 < iters < 2000
    ctx.apply_pressure(dt)
    b2, _ = ctx.download_rhs(dt)  # rhs of the projected field == -div/h
    assert np.abs(b2).max() < 5e-6
    ctx.close()


@pytest.mark.parametrize("scene", ["dam_break", "flip_obstacle", "pic_sphere"])
def test_fused_step_equals_staged_step_on_device(scene):
    """lfk_time_step (lean sort: velocity / c rows read through the permutation, fused advect+collide and
    correct+collide, gravity folded into P2G) against the same step issued stage by stage through the C ABI."""
    orc, rec = load_golden(scene)
    outs = []
    for fused in (True, False):
        ctx = DL.context_for(rec, max_iterations=2000)
        ctx.set_tuning("warm_start", 0)  # same initial guess (p = 0) on both sides: the comparison is to rounding
        ctx.upload_cells(rec["hash0/cells"])
        ctx.upload_particles(rec["collide1/particles"])  # old_position == position, as after a completed step
        for step in range(3):
            dt = 0.004 + 0.001 * step
            if fused:
                ctx.time_step(dt)
            else:
                ctx.advect(dt)
                ctx.collide()
                ctx.hash()
                ctx.p2g()
                ctx.gravity(dt)
                ctx.pressure_solve(dt)
                ctx.apply_pressure(dt)
                ctx.correct(dt)
                ctx.collide()
                ctx.extrapolate()
                ctx.g2p()
        outs.append((ctx.download_particles().copy(), ctx.download_cells().copy()))
        ctx.close()
    (pa, ca), (pb, cb) = outs
    assert np.array_equal(pa["raw_cell_index"], pb["raw_cell_index"])
    for f in ("position", "velocity", "cx", "cy", "cz"):
        assert PL.rel_l2(pa[f], pb[f]) < 1e-13, f
    assert np.array_equal(ca["type"], cb["type"])
    assert PL.rel_l2(ca["vel"], cb["vel"]) < 1e-13


@pytest.mark.parametrize("method", [capi.APIC, capi.FLIP, capi.PIC])
def test_cfl_after_a_step_is_the_exact_maximum_and_tracks_later_edits(method):
    """cfl() after a fused or staged step comes from the maximum the G2P kernel folded in; it must be the exact
    h / sqrt(max |v|^2) of the particles as downloaded, and anything that rewrites the particles afterwards (upload,
    seeding) must be seen by the next cfl()."""
    def expect(p):
        v = p["velocity"]
        s = v[:, 0] * v[:, 0]
        s = s + v[:, 1] * v[:, 1]
        s = s + v[:, 2] * v[:, 2]
        return 1.0 / np.sqrt(s.max())

    ctx = _device_scene(24, method)
    for step in range(3):
        ctx.time_step()
        assert ctx.cfl() == expect(ctx.download_particles())
    ctx.hash()  # a staged G2P on the sorted state refreshes the cache as well
    ctx.g2p()
    assert ctx.cfl() == expect(ctx.download_particles())
    p = ctx.download_particles().copy()
    p["velocity"] *= 0.25
    p["velocity"][7] = (1e4, -2e4, 3e4)
    ctx.upload_particles(p)
    assert ctx.cfl() == expect(p)
    ctx.time_step()
    ctx.seed_box_device((1.0, 20.0, 1.0), (2.0, 2.0, 2.0), velocity=(5e5, 0.0, 0.0), density=2, seed=5, append=True)
    assert ctx.cfl() == 1.0 / 5e5
    ctx.close()


def _device_scene(n=40, method=capi.APIC, **kw):
    """a sloshing block that fills ~half of an n^3 box, stepped a few times so that cells hold 0..20 particles"""
    ctx = capi.Context((n, n, n), cell_size=1.0, gravity=(0.0, -981.0, 0.0), method=method, max_iterations=2000, **kw)
    ctx.seed_box_device((0.0, 0.0, 0.0), (n * 0.55, n * 0.8, float(n)), density=2, seed=11)
    return ctx


@pytest.mark.parametrize("method", [capi.APIC, capi.FLIP, capi.PIC])
def test_p2g_kernel_variants_agree(method):
    """the z-marching P2G kernel (production) against the plain gather kernel (the reference's per-cell loop) on the
    same sorted state, through the staged call (velocity rows in place) and inside the fused step (lean sort: rows
    read through the permutation); cell types exact, faces to summation-order rounding"""
    ctx = _device_scene(method=method, blending_factor=0.95)
    for _ in range(4):
        ctx.time_step()
    ctx.hash()
    outs = {}
    for name, v in (("march", 0), ("gather", 2)):
        ctx.set_tuning("p2g", v)
        ctx.p2g()
        outs[name] = ctx.download_cells().copy()
        if method == capi.FLIP:
            outs[name + "/old"] = ctx.download_old_cells().copy()
    for other in ("gather",):
        assert np.array_equal(outs["march"]["type"], outs[other]["type"])
        assert PL.rel_l2(outs["march"]["vel"], outs[other]["vel"]) < 1e-13, other
        if method == capi.FLIP:
            assert PL.rel_l2(outs["march/old"]["vel"], outs[other + "/old"]["vel"]) < 1e-13, other
    assert np.abs(outs["march"]["vel"]).max() > 1.0  # a moving fluid, not an empty comparison
    # fused steps: march against gather from the same state
    parts, cells = ctx.download_particles().copy(), outs["march"]
    res = []
    for v in (0, 2):
        ctx.set_tuning("p2g", v)
        ctx.set_tuning("warm_start", 0)
        ctx.upload_cells(cells)
        ctx.upload_particles(parts)
        for _ in range(2):
            ctx.time_step(0.002)
        res.append((ctx.download_particles().copy(), ctx.download_cells().copy()))
    (pa, ca), (pb, cb) = res
    assert np.array_equal(pa["raw_cell_index"], pb["raw_cell_index"])
    assert np.array_equal(ca["type"], cb["type"])
    for f in ("position", "velocity", "cx", "cy", "cz"):
        assert PL.rel_l2(pa[f], pb[f]) < 1e-9, f  # two solves in between: agreement to solver tolerance
    ctx.close()


def test_warm_started_solve_reaches_the_same_tolerance():
    """fused step with the previous pressure as initial guess against the reference's p = 0 start: same residual
    tolerance, states agree to solver accuracy, and the warm start does not cost iterations"""
    res = []
    for warm in (1, 0):
        ctx = _device_scene()
        ctx.set_tuning("warm_start", warm)
        its = 0
        for _ in range(6):
            ctx.time_step(0.002)
            st = ctx.stats()
            assert st["pcg_residual"] < 1e-6
            its += st["pcg_iterations"]
        res.append((ctx.download_particles().copy(), its))
        ctx.close()
    (pa, ia), (pb, ib) = res
    assert np.array_equal(pa["raw_cell_index"], pb["raw_cell_index"])
    assert PL.rel_l2(pa["position"], pb["position"]) < 1e-8
    assert PL.rel_l2(pa["velocity"], pb["velocity"]) < 1e-4
    assert ia <= ib + 6, (ia, ib)


@pytest.mark.parametrize("method", [capi.FLIP, capi.PIC])
def test_untouched_payload_travels_with_its_particle(method):
    """PIC / FLIP never write the APIC c rows, so they are carried payload: after sorts (lean and full), fused and
    staged steps every particle must still hold the c rows it was uploaded with (tagged with an id and its start)"""
    ctx = _device_scene(n=24, method=method, blending_factor=0.95)
    parts = ctx.download_particles().copy()
    parts["cx"][:, 0] = np.arange(parts.shape[0])
    parts["cy"] = parts["position"]
    ctx.upload_particles(parts)
    for step in range(4):
        if step == 2:  # one step through the staged API (full sort)
            dt = 0.002
            ctx.advect(dt); ctx.collide(); ctx.hash(); ctx.p2g(); ctx.gravity(dt); ctx.pressure_solve(dt)
            ctx.apply_pressure(dt); ctx.correct(dt); ctx.collide(); ctx.extrapolate(); ctx.g2p()
        else:
            ctx.time_step(0.002)
        out = ctx.download_particles()
        ids = np.rint(out["cx"][:, 0]).astype(np.int64)
        assert np.array_equal(np.sort(ids), np.arange(parts.shape[0])), step
        assert np.abs(out["position"] - out["cy"]).max() < 1.0, step
        assert np.array_equal(out["cy"], parts["position"][ids]), step
    ctx.close()


def test_chunked_transfers_async_positions_and_checkpoint(tmp_path, monkeypatch):
    """N3: the chunked, double-buffered AoS transfers (forced to many chunks) round-trip bit for bit; the asynchronous
    positions download returns the positions of the moment it was queued even though the next step is issued before
    it is waited for; a checkpoint restored into a fresh context continues the run bit for bit (APIC and FLIP)."""
    monkeypatch.setenv("LFK_XFER_CHUNK", "777")
    for method in (capi.APIC, capi.FLIP):
        ctx = _device_scene(n=24, method=method, blending_factor=0.95)
        for _ in range(2):
            ctx.time_step()
        parts = ctx.download_particles().copy()
        assert parts.shape[0] > 10 * 777
        ctx.upload_particles(parts)
        back = ctx.download_particles()
        assert np.array_equal(back.view("u1"), parts.view("u1"))
        # asynchronous positions: queue, step, then wait
        n = ctx.num_particles()
        pin = capi.PinnedBuffer(n * 24)
        got = ctx.download_positions_async(pin.ptr.value, n)
        ctx.time_step(0.002)
        ctx.wait_transfers()
        xyz = pin.array[:n * 24].view(np.float64).reshape(n, 3)
        assert got == n and np.array_equal(xyz, parts["position"])
        pin.close()
        # checkpoint -> fresh context -> same continuation
        path = str(tmp_path / ("ckpt_%d.bin" % method))
        ctx.checkpoint_save(path)
        for _ in range(2):
            ctx.time_step(0.002)
        a, ca = ctx.download_particles().copy(), ctx.download_cells().copy()
        fresh = capi.Context((24, 24, 24), cell_size=1.0)
        fresh.checkpoint_load(path)
        assert int(fresh.params.method) == method and fresh.num_particles() == n
        for _ in range(2):
            fresh.time_step(0.002)
        b, cb = fresh.download_particles(), fresh.download_cells()
        assert np.array_equal(a.view("u1"), b.view("u1"))
        assert np.array_equal(ca["vel"], cb["vel"]) and np.array_equal(ca["type"], cb["type"])
        fresh.close()
        ctx.close()


@pytest.mark.skipif(not RB.available(), reason="oracle/_ref did not travel to this box")
def test_sources_on_device_match_reference():
    """N1: velocity coercion and seed_cell on the device against the reference's _advect_particles / _update_sources
    (src/simulation.cpp:227-238, 756-765, 136-151): the coerced + advected positions agree exactly, the spawned
    particles have the reference's per-cell counts and the source's velocity and lie inside their cell; two sources that
    share cells with different target densities follow the reference's count bookkeeping."""
    n = 16
    ref = RB.RefSim((n, n, n), method=RB.APIC)
    ref.seed_box((0, 0, 0), (0.5 * n, 0.4 * n, n))
    a_cells = [(x, y, z) for x in range(1, 4) for y in range(2, 9) for z in range(5, 9)]   # partly inside the water
    b_cells = [(x, y, z) for x in range(3, 5) for y in range(7, 10) for z in range(6, 8)]  # overlaps a_cells at x = 3
    ref.add_source(a_cells, (150.0, 0.0, -20.0), dens=2, coerce=True)
    ref.add_source(b_cells, (0.0, 60.0, 0.0), dens=3, coerce=False)
    ref.reset_space_hash()
    ref.time_step(0.004)  # a state with non-trivial velocities and c rows
    ctx = DL.context_for(ref, max_iterations=2000)
    ctx.set_sources([dict(cells=a_cells, velocity=(150.0, 0.0, -20.0), density=2, coerce=True),
                     dict(cells=b_cells, velocity=(0.0, 60.0, 0.0), density=3, coerce=False)], seed=11)
    dt = 0.003
    # ---- coercion + advection ----
    ref.update_and_hash()
    p0 = ref.particles()
    ctx.upload_cells(ref.cells())
    ctx.upload_particles(p0)
    ctx.coerce_sources()
    ctx.advect(dt)
    ref.advect(dt)
    a, b = ctx.download_particles(), ref.particles()
    for f in ("position", "velocity", "cx", "cy", "cz"):
        assert np.array_equal(a[f], b[f]), f
    assert (b["velocity"][:, 0] == 150.0).sum() > 50  # the scene does exercise coercion
    # ---- seed_cell ----
    ref.collide()
    ref.update_and_hash()  # (time_step re-keys the particles here, src/simulation.cpp:62; hash() alone would not)
    before = ref.particles()
    _, cnt_before = ref.space_hash()
    ctx.upload_particles(before)
    ctx.hash()
    added = ctx.update_sources()
    ref.update_sources()
    ref.update_and_hash()
    after_ref = ref.particles()
    assert added == after_ref.shape[0] - before.shape[0] and added > 0
    ctx.hash()
    _, cnt_dev = ctx.download_table()
    _, cnt_ref = ref.space_hash()
    assert np.array_equal(cnt_dev, cnt_ref)
    out = ctx.download_particles()
    old = set(map(bytes, np.ascontiguousarray(before["position"])))
    new = np.array([bytes(p) not in old for p in np.ascontiguousarray(out["position"])])
    assert new.sum() == added
    raw = out["raw_cell_index"][new]
    cell = np.stack([raw % n, (raw // n) % n, raw // (n * n)], axis=1)
    pos = out["position"][new]
    assert ((pos >= cell) & (pos < cell + 1)).all()
    in_a = np.array([tuple(c) in set(a_cells) for c in cell])
    va, vb = np.array([150.0, 0.0, -20.0]), np.array([0.0, 60.0, 0.0])
    vel = out["velocity"][new]
    assert ((vel == va).all(axis=1) | (vel == vb).all(axis=1)).all()
    assert (vel[~in_a] == vb).all()
    # same multiset of (cell, velocity) as the reference's spawned particles
    old_ref = np.array([bytes(p) not in old for p in np.ascontiguousarray(after_ref["position"])])
    key_dev = np.sort(raw * 4 + (vel[:, 0] == 150.0))
    key_ref = np.sort(after_ref["raw_cell_index"][old_ref] * 4 + (after_ref["velocity"][old_ref][:, 0] == 150.0))
    assert np.array_equal(key_dev, key_ref)
    # uniform in the cell: the mean in-cell fraction of a few hundred samples is close to 1/2
    assert np.abs((pos - cell).mean(axis=0) - 0.5).max() < 0.1
    # ---- the fused step with sources: counts follow the reference's ----
    ctx.upload_cells(ref.cells())
    ctx.upload_particles(after_ref)
    for step in range(3):
        ref.time_step(0.003)
        ctx.time_step(0.003)
        # the first step starts from identical states: identical counts; afterwards the spawned particles sit at
        # different (equally distributed) places, so the counts only track each other
        if step == 0:
            assert ctx.num_particles() == ref.num_particles()
        assert abs(ctx.num_particles() - ref.num_particles()) <= 0.02 * ref.num_particles()
    ctx.close()


def _mesh_box(lo, hi):
    (x0, y0, z0), (x1, y1, z1) = lo, hi
    v = np.array([[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0], [x0, y0, z1], [x1, y0, z1], [x1, y1, z1],
                  [x0, y1, z1]], dtype=np.float64)
    f = np.array([[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6], [0, 4, 5], [0, 5, 1], [1, 5, 6], [1, 6, 2], [2, 6, 7],
                  [2, 7, 3], [3, 7, 4], [3, 4, 0]], dtype=np.uint64)
    return v, f


def _mesh_sphere(center, radius, nu=24, nv=16):
    """a closed UV sphere (triangles only; skinny triangles at the poles exercise the separating-axis test)"""
    th = np.linspace(0.0, np.pi, nv + 1)[1:-1]
    ph = np.linspace(0.0, 2.0 * np.pi, nu, endpoint=False)
    ring = np.stack([np.outer(np.sin(th), np.cos(ph)), np.outer(np.cos(th), np.ones_like(ph)),
                     np.outer(np.sin(th), np.sin(ph))], axis=2).reshape(-1, 3)
    v = np.concatenate([[[0.0, 1.0, 0.0]], ring, [[0.0, -1.0, 0.0]]]) * radius + np.asarray(center)
    f = []
    for j in range(nu):
        f.append([0, 1 + j, 1 + (j + 1) % nu])
        for i in range(nv - 2):
            a, b = 1 + i * nu + j, 1 + i * nu + (j + 1) % nu
            f += [[a, a + nu, b], [b, a + nu, b + nu]]
        base = 1 + (nv - 2) * nu
        f.append([base + j, len(v) - 1, base + (j + 1) % nu])
    return v, np.array(f, dtype=np.uint64)


@pytest.mark.skipif(not RB.available() or not hasattr(RB.lib(), "ref_mesher_sample"),
                    reason="oracle/_ref (with the mesher / voxelizer sources) did not travel to this box")
def test_mesher_sampling_on_device_matches_reference():
    """N2: mesher::_sample_surface_function (src/mesher.cpp:333-376) on the device against the reference, on host
    positions with the testbed's settings (testbed/main.cpp:218-224) and on the context's own resident particles:
    the same points read 1.0 (nothing in range) and NaN (only zero weights), the rest agree to summation order."""
    ctx = _device_scene(n=24)
    for _ in range(3):
        ctx.time_step()
    pos = ctx.download_positions()
    for size, off, cs, ext, rad, r, own in (((56, 56, 56), (-1.0, -1.0, -1.0), 0.5, 2.0, 3, 0.5, False),
                                           ((30, 26, 22), (0.3, -0.7, 1.1), 0.9, 0.8, 2, 0.35, True),
                                           ((20, 20, 20), (2.0, 2.0, 2.0), 1.0, 0.5, 1, 0.5, True)):
        want = RB.mesher_sample(size, off, cs, ext, rad, pos, r)
        got = ctx.mesher_sample(size, off, cs, ext, rad, r, xyz=None if own else pos)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.array_equal(got == 1.0, want == 1.0)
        ok = ~np.isnan(want)
        assert (want[ok] != 1.0).sum() > 1000
        assert np.abs(got[ok] - want[ok]).max() <= 1e-12 * max(1.0, np.abs(want[ok]).max())
    ctx.close()


@pytest.mark.skipif(not RB.available() or not hasattr(RB.lib(), "ref_voxelize"),
                    reason="oracle/_ref (with the mesher / voxelizer sources) did not travel to this box")
def test_voxelizer_on_device_matches_reference():
    """N4: voxelizer (src/voxelizer.cpp:19-126) + obstacle (src/data_structures/obstacle.cpp:9-29) on the device: the
    voxel classification (interior / exterior / surface) is exact for boxes and spheres at several cell sizes and
    offsets, including a hollow shell (an interior that the flood must not reach) and a mesh whose bounding grid is one
    solid surface block; the obstacle cells equal the reference's wherever the reference's loop bounds are defined, and
    marking them solid changes exactly those cell types."""
    n = 20
    ctx = capi.Context((n, n, n), cell_size=1.0)
    shell_v0, shell_f0 = _mesh_box((2.2, 2.1, 2.3), (15.4, 14.9, 13.7))
    shell_v1, shell_f1 = _mesh_box((5.2, 5.1, 5.3), (11.4, 10.9, 9.7))
    shell = (np.concatenate([shell_v0, shell_v1]), np.concatenate([shell_f0, shell_f1 + 8]))
    cases = [(_mesh_box((1.3, 1.2, 1.1), (7.7, 6.5, 5.8)), 1.0, (0.0, 0.0, 0.0)),
             (_mesh_box((0.3, -0.8, 1.1), (7.7, 6.5, 5.8)), 1.0, (0.0, 0.0, 0.0)),
             (_mesh_box((3.3, 2.2, 4.1), (9.7, 6.5, 8.8)), 1.0, (0.0, 0.0, 0.0)),
             (_mesh_sphere((8.1, 7.4, 9.2), 5.3), 1.0, (0.0, 0.0, 0.0)),
             (_mesh_sphere((3.0, 2.5, 3.5), 2.4), 0.37, (-0.2, 0.1, 0.3)),
             (shell, 1.0, (0.0, 0.0, 0.0)),
             (_mesh_box((4.1, 4.2, 4.3), (4.6, 4.7, 4.8)), 1.0, (0.0, 0.0, 0.0))]
    defined = 0
    for (v, f), cs, off in cases:
        ctx.set_params(cell_size=cs, grid_offset=off)
        gmin_r, vox_r, cells_r = RB.voxelize(v, f, cs, off, (n, n, n))
        gmin_d, vox_d = ctx.voxelize_mesh(v, f, cs, off)
        assert np.array_equal(gmin_d, gmin_r) and vox_d.shape == vox_r.shape
        assert np.array_equal(vox_d, vox_r)
        cells_d = ctx.obstacle_cells()
        # the intended set: interior voxels inside the simulation grid, z / y / x ascending
        zz, yy, xx = np.nonzero(vox_r == 0)
        g = np.stack([xx, yy, zz], axis=1).astype(np.int64) + gmin_r
        g = g[((g >= 0) & (g < n)).all(axis=1)]
        assert np.array_equal(cells_d.astype(np.int64), g)
        if cells_r is not None and (gmin_r == 0).all():
            defined += 1
            assert np.array_equal(cells_d, cells_r)
        before = ctx.download_cells()["type"].copy()
        ctx.obstacle_cells(mark_solid=True)
        after = ctx.download_cells()["type"]
        raw = g[:, 0] + n * (g[:, 1] + n * g[:, 2])
        expect = before.copy()
        expect[raw] = capi.SOLID
        assert np.array_equal(after, expect)
        ctx.upload_cells(_air_cells(n))
    assert defined >= 1
    ctx.close()


def _air_cells(n):
    c = np.zeros(n ** 3, dtype=capi.CELL_DTYPE)
    c["type"] = capi.AIR
    return c


def test_two_gpu_slabs_match_single_gpu():
    """z-slab decomposition with particle migration, ghost copies and NCCL halos against the single-GPU run of the
    same scene (tests/mgpu_check.py under torchrun; needs >= 2 GPUs)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(os.path.dirname(__file__), "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
