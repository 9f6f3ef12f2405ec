"""Shared helpers for the oracle-pinning and parity tests (test infrastructure)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oraclebind as OB  # noqa: E402
from oracle import refbind as RB  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def make_scene(name, n=16):
    """The BASELINE configs, scaled to an n^3 grid; returns a seeded RefSim."""
    if name == "cube_drop":      # testbed setup 0 (testbed/main.cpp:138-140), APIC
        s = RB.RefSim((n, n, n), method=RB.APIC)
        s.seed_box((0.3 * n, 0.3 * n, 0.3 * n), (0.4 * n, 0.4 * n, 0.4 * n))
    elif name == "dam_break":    # testbed setup 3 (:148-150), APIC
        s = RB.RefSim((n, n, n), method=RB.APIC)
        s.seed_box((0, 0, 0), (0.2 * n, n, n))
    elif name == "flip_obstacle":  # config 2: FLIP 0.95, solid box [2n/5,3n/5) x [0,n/4) x [2n/5,3n/5)
        s = RB.RefSim((n, n, n), method=RB.FLIP, blend=0.95)
        m = np.zeros((n, n, n), dtype=bool)
        m[2 * n // 5:3 * n // 5, 0:n // 4, 2 * n // 5:3 * n // 5] = True  # [z, y, x]
        s.set_solid(m)
        s.seed_box((0, 0, 0), (0.3 * n, 0.8 * n, 0.3 * n))
    elif name == "pic_sphere":   # testbed setup 1 (:141-143), PIC, non-unit h and offset, non-cubic grid
        s = RB.RefSim((n, n + 3, n - 2), h=0.5, offset=(-1.0, 0.25, 2.0), method=RB.PIC)
        s.seed_sphere((-1.0 + 0.25 * n, 0.25 + 0.3 * n, 2.0 + 0.22 * n), 0.15 * n)
    else:
        raise KeyError(name)
    return s


SCENES = ("cube_drop", "dam_break", "flip_obstacle", "pic_sphere")


def make_random_scene(seed):
    """A randomised small scene: non-cubic grid, cell size that need not be a power of two (true divisions; APIC's
    weights without /h and face positions built by repeated addition), offset grid, any
    of the three methods, random solid blocks, one or two moving blocks / spheres of water.  Returns (RefSim, rng);
    None when the draw seeded no particle."""
    rng = np.random.default_rng(1000 + seed)
    size = tuple(int(v) for v in rng.integers(6, 15, size=3))  # (nx, ny, nz)
    h = float(rng.choice([0.25, 0.5, 1.0, 1.7, 2.0, 0.3]))
    offset = tuple(float(v) for v in rng.uniform(-3.0, 3.0, size=3))
    method = int(rng.choice([RB.APIC, RB.FLIP, RB.PIC]))
    blend = float(rng.choice([1.0, 0.95, 0.5]))
    ref = RB.RefSim(size, h=h, offset=offset, gravity=(0.0, -981.0 * h, 0.0), method=method, blend=blend)
    nx, ny, nz = size
    solid = np.zeros((nz, ny, nx), dtype=bool)
    for _ in range(int(rng.integers(0, 3))):
        lo = [int(rng.integers(0, d - 1)) for d in (nz, ny, nx)]
        ext = [int(rng.integers(1, 4)) for _ in range(3)]
        solid[lo[0]:lo[0] + ext[0], lo[1]:lo[1] + ext[1], lo[2]:lo[2] + ext[2]] = True
    if solid.any() and not solid.all():
        ref.set_solid(solid)
    ext = np.array(size, dtype=np.float64) * h
    off = np.array(offset)
    for _ in range(int(rng.integers(1, 3))):
        a = off + rng.uniform(0.0, 0.5, size=3) * ext
        b = rng.uniform(0.2, 0.5, size=3) * ext
        vel = tuple(float(v) for v in rng.uniform(-40.0, 40.0, size=3) * h)
        if rng.random() < 0.3:
            ref.seed_sphere(tuple(a + 0.5 * b), float(0.4 * b.min()), vel=vel)
        else:
            ref.seed_box(tuple(a), tuple(b), vel=vel)
    if ref.num_particles() == 0:
        return None, rng
    ref.reset_space_hash()
    return ref, rng


def oracle_for(ref):
    """A C-restatement Oracle configured like the given RefSim."""
    return OB.Oracle(ref.size, h=ref.h, offset=ref.offset, gravity=ref.gravity, method=ref.method,
                     blend=ref.blend, density=ref.density, skin=ref.skin, stiffness=ref.stiffness,
                     extrap_iters=ref.extrap_iters)


def split_particles(arr):
    """RefSim AoS records -> contiguous SoA-ish arrays used by the oracle."""
    pos = np.ascontiguousarray(arr["position"])
    vel = np.ascontiguousarray(arr["velocity"])
    c = np.ascontiguousarray(np.concatenate([arr["cx"], arr["cy"], arr["cz"]], axis=1))
    old = np.ascontiguousarray(arr["old_position"])
    key = np.ascontiguousarray(arr["raw_cell_index"])
    return pos, vel, c, old, key


def split_cells(arr):
    return np.ascontiguousarray(arr["vel"]), np.ascontiguousarray(arr["type"])


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    d = np.linalg.norm(a - b)
    nb = np.linalg.norm(b)
    return d / nb if nb > 0 else d


def degenerate_mask(pos):
    """True for particles that have another particle within sqrt(1e-12) (the reference's random-kick branch)."""
    from scipy.spatial import cKDTree
    pairs = cKDTree(pos).query_pairs(1.0001e-6, output_type="ndarray")
    m = np.zeros(pos.shape[0], dtype=bool)
    m[pairs.ravel()] = True
    return m


PHASES = ("hash0", "advect", "collide1", "hash", "p2g", "gravity", "solve", "apply_pressure", "correct",
          "collide2", "extrapolate", "g2p")


def record_step(ref, dt):
    """Replays one reference time_step and returns {name: array} with the state after every phase."""
    rec = {}

    def hook(name, r):
        if name in ("hash0", "advect", "collide1", "hash", "correct", "collide2", "g2p"):
            rec[name + "/particles"] = r.particles()
        if name in ("hash0", "hash", "p2g", "gravity", "apply_pressure", "extrapolate"):
            rec[name + "/cells"] = r.cells()
        if r.method == RB.FLIP and name in ("p2g", "gravity"):
            rec[name + "/old_cells"] = r.old_cells()
        if name == "hash":
            b, c = r.space_hash()
            rec["hash/begin"], rec["hash/count"] = b, c
            rec["hash/fluid_cells"] = r.fluid_cells()
        if name == "gravity":
            r.solver_setup()
            bb, fl = r.solver_rhs(dt)
            rec["rhs/b"], rec["rhs/flags"] = bb, fl

    p, res, it = ref.replay_time_step(dt, hook)
    rec["solve/p"] = p
    rec["solve/residual"] = np.float64(res)
    rec["solve/iters"] = np.uint64(it)
    rec["dt"] = np.float64(dt)
    rec["g2p/cfl"] = np.float64(ref.cfl())
    return rec


def check_oracle_against_record(orc, rec, exact=True):
    """Runs every stage of the C restatement on the recorded inputs and compares with the recorded outputs.
    Returns the list of mismatching stage names (bit-exact comparison)."""
    dt = float(rec["dt"])
    flip = orc.P.method == OB.FLIP
    bad = []

    def chk(name, ok):
        if not ok:
            bad.append(name)

    p0, v0, c0, o0, _ = split_particles(rec["hash0/particles"])
    _, ty0 = split_cells(rec["hash0/cells"])
    chk("keys0", np.array_equal(orc.cell_keys(p0), rec["hash0/particles"]["raw_cell_index"]))
    pa = p0.copy()
    orc.advect(dt, pa, v0)
    chk("advect", np.array_equal(pa, rec["advect/particles"]["position"]))
    pc = pa.copy()
    orc.collide(pc, np.ascontiguousarray(rec["advect/particles"]["old_position"]), ty0)
    chk("collide1", np.array_equal(pc, rec["collide1/particles"]["position"]))
    # sort: same multiset per cell, same table (the reference's in-cell order is unspecified)
    ph, vh, ch, _, keyh = split_particles(rec["hash/particles"])
    key = orc.cell_keys(pc)
    perm, begin, count, fluid = orc.hash(key)
    chk("sorted_keys", np.array_equal(key[perm], keyh))
    chk("table", np.array_equal(count, rec["hash/count"]) and np.array_equal(fluid, rec["hash/fluid_cells"]) and
        np.array_equal(begin[count > 0], rec["hash/begin"][count > 0]))
    b2, c2, f2 = orc.cell_ranges(keyh)
    chk("ranges", np.array_equal(b2, rec["hash/begin"]) and np.array_equal(c2, rec["hash/count"]) and
        np.array_equal(f2, rec["hash/fluid_cells"]))
    # P2G in the reference's own in-cell order => bit-exact
    gv, ty = split_cells(rec["hash/cells"])
    gv, ty, ogv = gv.copy(), ty.copy(), np.zeros_like(gv)
    orc.p2g(ph, vh, ch, b2, c2, gv, ty, ogv)
    rgv, rty = split_cells(rec["p2g/cells"])
    chk("p2g", np.array_equal(gv, rgv) and np.array_equal(ty, rty))
    if flip:
        chk("p2g_old", np.array_equal(ogv, rec["p2g/old_cells"]["vel"]))
    orc.gravity(dt, gv)
    chk("gravity", np.array_equal(gv, rec["gravity/cells"]["vel"]))
    imap, flags, b = orc.solver_setup(gv, ty, f2)
    chk("rhs", np.array_equal(b, rec["rhs/b"]) and np.array_equal(flags, rec["rhs/flags"]))
    p, res, it = orc.solve(dt, f2, imap, flags, b)
    chk("solve", np.array_equal(p, rec["solve/p"]) and res == float(rec["solve/residual"]) and
        it == int(rec["solve/iters"]))
    orc.apply_pressure(dt, f2, imap, p, gv, ty)
    chk("apply_pressure", np.array_equal(gv, rec["apply_pressure/cells"]["vel"]))
    pcor = ph.copy()
    orc.correct(dt, pcor, b2, c2)
    keep = ~degenerate_mask(ph)  # r^2 < 1e-12 pairs take a std::random_device kick in the reference (:584-587)
    chk("correct", np.array_equal(pcor[keep], rec["correct/particles"]["position"][keep]))
    pc2 = np.ascontiguousarray(rec["correct/particles"]["position"]).copy()
    orc.collide(pc2, np.ascontiguousarray(rec["correct/particles"]["old_position"]), ty)
    chk("collide2", np.array_equal(pc2, rec["collide2/particles"]["position"]))
    orc.extrapolate(f2, gv, ty)
    chk("extrapolate", np.array_equal(gv, rec["extrapolate/cells"]["vel"]))
    vg, cg = vh.copy(), ch.copy()
    orc.g2p(pc2, vg, cg, gv, ogv)
    _, vr, cr, _, _ = split_particles(rec["g2p/particles"])
    chk("g2p", np.array_equal(vg, vr) and np.array_equal(cg, cr))
    chk("cfl", orc.cfl(vg) == float(rec["g2p/cfl"]))
    return bad
