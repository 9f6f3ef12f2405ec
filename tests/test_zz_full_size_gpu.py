"""Full-size property test (runs last: the file name sorts after the other test modules, so a failure here cannot
hide their results under `pytest -x`).  Needs a GPU with >= 40 GB."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libfluid_b200 import capi  # noqa: E402

pytestmark = pytest.mark.gpu


def test_full_size_step_properties_256():
    """BASELINE configs[2] at its full size (256^3 APIC, 130 M particles; the bench scene): size-independent
    properties instead of an oracle run -- particle count conserved, solver residual below the reference's
    tolerance, the sorted particle array is ordered by the reference's cell key and agrees with the {begin, count}
    table, and after P2G + gravity + solve + pressure update the discrete divergence of every fluid cell is below
    the tolerance."""
    import bench as B
    n = 256
    ctx = capi.Context((n, n, n), cell_size=1.0, gravity=B.GRAVITY, method=capi.APIC, max_iterations=1000)
    for k, (start, size) in enumerate(B.scene_boxes(n, n, n)):
        ctx.seed_box_device(start, size, density=2, seed=20261017, append=k > 0)
    np0 = ctx.num_particles()
    assert np0 > 120_000_000
    for _ in range(2):
        ctx.time_step()
        st = ctx.stats()
        assert st["pcg_residual"] < 1e-6
    assert ctx.num_particles() == np0
    ctx.hash()
    begin, count = ctx.download_table()
    assert int(count.sum()) == np0
    filled = count > 0  # the reference leaves begin = 0 in empty cells (reset_space_hash, src/simulation.cpp:131-134)
    assert np.array_equal(begin[filled], (np.cumsum(count) - count)[filled]) and not begin[~filled].any()
    pos = ctx.download_positions()
    assert np.isfinite(pos).all() and pos.min() >= 0.0 and pos.max() <= float(n)
    key = np.minimum(pos[:, 2].astype(np.int64), n - 1)  # src/simulation.cpp:251-261 with h = 1, offset 0
    for d in (1, 0):
        key *= n
        key += np.minimum(pos[:, d].astype(np.int64), n - 1)
    del pos
    assert bool((key[1:] >= key[:-1]).all())
    assert np.array_equal(np.bincount(key, minlength=n ** 3).astype(np.uint64), count)
    del key
    dt = 0.002
    ctx.p2g()
    ctx.gravity(dt)
    res, iters = ctx.pressure_solve(dt)
    assert res < 1e-6 and 0 < iters < 1000
    ctx.apply_pressure(dt)
    b2, _ = ctx.download_rhs(dt)  # rhs of the projected field == -div/h
    assert np.abs(b2).max() < 5e-6
    ctx.close()


@pytest.mark.parametrize("seed", [0, 3, 4, 5])
def test_stages_match_reference_randomised(seed):
    """Every stage of the CUDA path against the compiled reference on randomised small scenes
    (pinlib.make_random_scene): non-cubic and offset grids, cell sizes 1.7 and 0.3 (not powers of two: the true-division
    paths), 0.5 and 0.25, APIC / FLIP / PIC, solid blocks, water thrown at walls.  Same checks and tolerances as
    test_stages_match_reference_live."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import devlib as DL
    import pinlib as PL
    if not PL.RB.available():
        pytest.skip("oracle/_ref did not travel to this box")
    ref, rng = PL.make_random_scene(seed)
    assert ref is not None
    orc = PL.oracle_for(ref)
    ctx = DL.context_for(ref, max_iterations=2000)
    for step in range(6):
        dt = min(ref.cfl_number * ref.cfl(), 0.033) if step % 2 else float(rng.choice([0.002, 0.006]))
        rec = PL.record_step(ref, dt)
        assert DL.check_device_against_record(ctx, rec, orc) == {}, "seed %d step %d" % (seed, step)
    ctx.close()
