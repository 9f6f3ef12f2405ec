"""Parity at the sizes BASELINE.json names, and across the kernels' tile seams -- every stage of the CUDA path against
the compiled reference driven live (oracle/_ref travels to the GPU box; nothing here reads /root/reference).

  * configs[0]: 64^3 dam break (testbed setup 3 scaled by 64/50), APIC, 4.2e5 particles, three
    consecutive recorded steps (all of them with PCG iterations);
  * configs[1]: 128^3 column collapse with the solid box, FLIP 0.95, 1.2e6 particles, one recorded step after two
    plain ones;
  * a 48 x 20 x 12 scene whose water crosses x = 16, 30 and 32 (the position-correction tile width is 16, the P2G block
    width 30) with motion along x;
  * a scene seeded at 27 particles per cell, so that the position correction's staged tile overflows and every
    particle takes its global-memory path.
Tolerances: those of devlib.check_device_against_record.
"""
import numpy as np
import pytest

import devlib as DL
import pinlib as PL
from pinlib import RB

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not RB.available(), reason="oracle/_ref did not travel to this box")]


def _run(ref, steps, plain_first=0, max_iterations=4000):
    orc = PL.oracle_for(ref)
    ctx = DL.context_for(ref, max_iterations=max_iterations)
    for _ in range(plain_first):
        ref.time_step(min(ref.cfl_number * ref.cfl(), 0.033))
    its = []
    for step in range(steps):
        dt = min(ref.cfl_number * ref.cfl(), 0.033)
        rec = PL.record_step(ref, dt)
        its.append(int(rec["solve/iters"]))
        assert DL.check_device_against_record(ctx, rec, orc) == {}, "step %d" % step
    ctx.close()
    return its


def test_config0_dam_break_64_apic():
    n = 64
    ref = RB.RefSim((n, n, n), method=RB.APIC)
    ref.seed_box((0, 0, 0), (0.2 * n, n, n))
    ref.reset_space_hash()
    assert ref.num_particles() > 400_000
    its = _run(ref, 3)
    assert max(its) > 0


def test_config1_column_collapse_128_flip_obstacle():
    n = 128
    ref = RB.RefSim((n, n, n), method=RB.FLIP, blend=0.95)
    m = np.zeros((n, n, n), dtype=bool)
    m[2 * n // 5:3 * n // 5, 0:n // 4, 2 * n // 5:3 * n // 5] = True  # [z, y, x]
    ref.set_solid(m)
    ref.seed_box((0, 0, 0), (0.3 * n, 0.8 * n, 0.3 * n))
    ref.reset_space_hash()
    assert ref.num_particles() > 1_000_000
    its = _run(ref, 1, plain_first=2)
    assert its[0] > 0


@pytest.mark.parametrize("method", [RB.APIC, RB.FLIP])
def test_water_across_the_x_tile_seams(method):
    ref = RB.RefSim((48, 20, 12), method=method, blend=0.95)
    ref.seed_box((8.0, 0.0, 0.0), (37.5, 11.0, 12.0), vel=(35.0, 0.0, 4.0))
    ref.reset_space_hash()
    its = _run(ref, 4)
    assert max(its) > 0


def test_overfull_tiles_take_the_global_path():
    ref = RB.RefSim((40, 8, 8), method=RB.APIC)
    ref.seed_box((1.0, 0.0, 0.0), (36.0, 5.0, 8.0), vel=(10.0, 0.0, 0.0), dens=3)
    ref.reset_space_hash()
    assert ref.num_particles() > 27 * 35 * 4 * 8
    _run(ref, 2)
