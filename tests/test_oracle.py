"""CPU tests of the test infrastructure itself: the plain-C restatement (oracle/fluid_oracle.c) must agree
bit for bit with (a) the golden vectors minted from the unmodified reference and (b), where oracle/_ref was
built, the compiled reference driven live -- stage by stage, over several steps and all three methods."""
import os

import numpy as np
import pytest

import pinlib as PL
from pinlib import OB, RB


def load_golden(scene):
    z = np.load(os.path.join(PL.GOLDEN_DIR, scene + ".npz"))
    rec = {k: z[k] for k in z.files}
    orc = OB.Oracle(rec["meta/size"], h=float(rec["meta/h"]), offset=rec["meta/offset"],
                    gravity=rec["meta/gravity"], method=int(rec["meta/method"]), blend=float(rec["meta/blend"]))
    return orc, rec


@pytest.mark.parametrize("scene", PL.SCENES)
def test_oracle_matches_golden(scene):
    orc, rec = load_golden(scene)
    assert int(rec["solve/iters"]) > 0  # the recorded step exercises the PCG loop
    assert PL.check_oracle_against_record(orc, rec) == []


@pytest.mark.skipif(not RB.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_replay_is_bit_identical_to_stock_time_step():
    a, b = PL.make_scene("dam_break", 12), PL.make_scene("dam_break", 12)
    for _ in range(5):
        a.time_step(0.005)
        b.replay_time_step(0.005)
    assert np.array_equal(a.particles().view("u1"), b.particles().view("u1"))
    assert np.array_equal(a.cells()["vel"], b.cells()["vel"])


@pytest.mark.skipif(not RB.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("scene", PL.SCENES)
def test_oracle_pinned_to_reference_live(scene):
    ref = PL.make_scene(scene, 14)
    orc = PL.oracle_for(ref)
    iters = 0
    for step in range(30):
        dt = min(ref.cfl_number * ref.cfl(), 0.033) if step % 2 else 0.004  # time_step(): src/simulation.cpp:127-129
        rec = PL.record_step(ref, dt)
        iters += int(rec["solve/iters"])
        assert PL.check_oracle_against_record(orc, rec) == [], "step %d" % step
    assert iters > 0


def test_known_answers():
    """Hand-derivable cases from SURVEY.md 8(c)."""
    n = 6
    orc = OB.Oracle((n, n, n), gravity=(0, 0, 0))
    # (iii) one particle at a cell centre: its velocity lands on the 6 nearest faces with weight 1/2
    pos = np.array([[2.5, 3.5, 1.5]])
    vel = np.array([[1.0, -2.0, 3.0]])
    c = np.zeros((1, 9))
    key = orc.cell_keys(pos)
    assert key[0] == 2 + n * (3 + n * 1)
    perm, begin, count, fluid = orc.hash(key)
    gv = np.zeros((n ** 3, 3))
    ty = np.full(n ** 3, OB.AIR, dtype=np.uint8)
    orc.p2g(pos, vel, c, begin, count, gv, ty)
    me = int(key[0])
    assert ty[me] == OB.FLUID and (ty == OB.FLUID).sum() == 1
    assert gv[me, 0] == 1.0 and gv[me - 1, 0] == 1.0
    assert gv[me, 1] == -2.0 and gv[me - n, 1] == -2.0
    assert gv[me, 2] == 3.0 and gv[me - n * n, 2] == 3.0
    # (i) divergence-free uniform flow => b == 0 => early out (p = 0, residual 0, 0 iterations)
    pos = (np.indices((2, 2, 2)).reshape(3, -1).T + 2.25).astype(np.float64)
    pos = np.concatenate([pos, pos + 0.5])
    vel = np.tile([[0.0, -3.0, 0.0]], (pos.shape[0], 1))
    key = orc.cell_keys(pos)
    perm, begin, count, fluid = orc.hash(key)
    gv[:] = 0
    ty[:] = OB.AIR
    orc.p2g(pos[perm], vel[perm], np.zeros((pos.shape[0], 9)), begin, count, gv, ty)
    imap, flags, b = orc.solver_setup(gv, ty, fluid)
    assert np.abs(b).max() < 1e-12
    p, res, it = orc.solve(0.01, fluid, imap, flags, b)
    assert it == 0 and res == 0.0 and not p.any()


def test_empty_and_degenerate_inputs():
    orc = OB.Oracle((4, 5, 3))
    e3, e9 = np.zeros((0, 3)), np.zeros((0, 9))
    key = orc.cell_keys(e3)
    perm, begin, count, fluid = orc.hash(key)
    assert fluid.size == 0 and not count.any()
    gv = np.ones((60, 3))
    ty = np.full(60, OB.FLUID, dtype=np.uint8)
    ty[7] = OB.SOLID
    orc.p2g(e3, e3, e9, begin, count, gv, ty)
    assert not gv.any() and ty[7] == OB.SOLID and (np.delete(ty, 7) == OB.AIR).all()
    assert orc.cfl(e3) == np.inf
    # positions outside the grid clamp into it (src/simulation.cpp:255-257)
    pos = np.array([[-3.0, 2.0, 1.0], [100.0, 100.0, 100.0], [3.999999, 4.999999, 2.999999]])
    assert list(orc.cell_keys(pos)) == [0 + 4 * (2 + 5 * 1), 59, 59]


@pytest.mark.skipif(not RB.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", range(10))
def test_oracle_pinned_to_reference_randomised(seed):
    """Randomised pinning (PL.make_random_scene: non-cubic grids, cell sizes that are not powers of two, offset
    grids, all three methods, solid blocks, moving water) -- every stage of every step bit for bit against the
    compiled reference."""
    ref, rng = PL.make_random_scene(seed)
    if ref is None:
        pytest.skip("empty scene")
    orc = PL.oracle_for(ref)
    for step in range(6):
        dt = min(ref.cfl_number * ref.cfl(), 0.033) if step % 2 else float(rng.choice([0.002, 0.006]))
        rec = PL.record_step(ref, dt)
        assert PL.check_oracle_against_record(orc, rec) == [], "seed %d step %d" % (seed, step)
