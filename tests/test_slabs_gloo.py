"""World-size-2 (and 3) CPU tests of the multi-GPU rank protocol over gloo: the slab partition, the particle
exchange rule (who sends what to whom, who keeps a ghost copy, who drops) and the bookkeeping identities the CUDA
path relies on, with numpy standing in for the device arrays.  The CUDA kernels themselves are covered by
tests/mgpu_check.py on real GPUs; nothing here touches the oracle or computes any physics."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from libfluid_b200 import slabs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _xchg(rank, world, send_up, send_dn):
    """neighbour exchange of variable-size float64 arrays: counts first, then payload (the lfkx protocol)"""
    got = {}
    for peer, payload, key in ((rank + 1, send_up, "up"), (rank - 1, send_dn, "dn")):
        if not 0 <= peer < world:
            got[key] = np.zeros((0, 3))
            continue
        cnt = torch.tensor([payload.shape[0]], dtype=torch.int64)
        rcnt = torch.zeros(1, dtype=torch.int64)
        reqs = [dist.isend(cnt, peer), dist.irecv(rcnt, peer)]
        [r.wait() for r in reqs]
        buf = torch.zeros((int(rcnt.item()), 3), dtype=torch.float64)
        reqs = [dist.isend(torch.from_numpy(np.ascontiguousarray(payload)), peer), dist.irecv(buf, peer)]
        [r.wait() for r in reqs]
        got[key] = buf.numpy()
    return got


def _worker(rank, world, port, nz, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z0, nzl = slabs.slab_range(nz, world, rank)
        # 1. the slabs tile [0, nz) in rank order
        all_ranges = [None] * world
        dist.all_gather_object(all_ranges, (z0, nzl))
        assert all_ranges[0][0] == 0 and sum(r[1] for r in all_ranges) == nz
        assert all(all_ranges[k][0] + all_ranges[k][1] == all_ranges[k + 1][0] for k in range(world - 1))
        assert (slabs.owner_of_z(nz, world, np.arange(z0, z0 + nzl)) == rank).all()
        # 2. own particles, then one "advection" of at most 3 cells (cfl_number) along z
        rng = np.random.default_rng(100 + rank)
        p = rng.uniform([0, 0, z0], [8, 8, z0 + nzl], size=(4000, 3))
        p[:, 2] = np.clip(p[:, 2] + rng.uniform(-3, 3, size=p.shape[0]), 0.01, nz - 0.01)
        zc = slabs.z_cell(p[:, 2], nz)
        up, dn, dead = slabs.classify(zc, z0, nzl, rank + 1 < world, rank > 0)
        assert not (up & dn).any()                      # slabs >= 4 thick: never both
        assert (dead <= (up | dn)).all()                # whatever is dropped here was sent to someone
        got = _xchg(rank, world, p[up], p[dn])
        merged = np.concatenate([p[~dead], got["dn"], got["up"]])
        zm = slabs.z_cell(merged[:, 2], nz)
        glo, own, ghi = slabs.split_after_sort(zm, z0, nzl)
        assert (glo | own | ghi).all()                  # nothing lands outside slab + ghost layers
        # 3. conservation and ghost consistency across ranks
        counts = [None] * world
        dist.all_gather_object(counts, (int(own.sum()), p.shape[0]))
        assert sum(c[0] for c in counts) == sum(c[1] for c in counts)
        mine_top = merged[own & (zm == z0 + nzl - 1)]
        mine_bot = merged[own & (zm == z0)]
        ghosts = [None] * world
        dist.all_gather_object(ghosts, (merged[glo], merged[ghi], mine_bot, mine_top))

        def same(a, b):
            return a.shape == b.shape and np.array_equal(a[np.lexsort(a.T)], b[np.lexsort(b.T)])
        if rank + 1 < world:  # my upper ghost layer == the upper rank's bottom own layer
            assert same(merged[ghi], ghosts[rank + 1][2])
        if rank > 0:
            assert same(merged[glo], ghosts[rank - 1][3])
        ret[rank] = "ok"
    except BaseException as ex:  # noqa: BLE001
        ret[rank] = "%s: %s" % (type(ex).__name__, ex)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nz", [(2, 16), (2, 11), (3, 14)])
def test_exchange_protocol_over_gloo(world, nz):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nz, ret)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(not p.is_alive() for p in procs)
    assert dict(ret) == {r: "ok" for r in range(world)}


def test_slab_rules():
    assert slabs.slab_range(256, 1, 0) == (0, 256)
    assert [slabs.slab_range(19, 4, r) for r in range(4)] == [(0, 5), (5, 5), (10, 5), (15, 4)]
    with pytest.raises(ValueError):
        slabs.slab_range(15, 4, 0)
    zc = np.array([3, 4, 5, 8, 9, 10, 11])
    up, dn, dead = slabs.classify(zc, 5, 5, True, True)  # slab [5, 10)
    assert list(up) == [False, False, False, False, True, True, True]
    assert list(dn) == [True, True, True, False, False, False, False]
    assert list(dead) == [True, False, False, False, False, False, True]


def _agree_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = []
        # bench.py's end-to-end loop: the particle count of a rank drifts from step to step; the ranks agree on every
        # local capacity check BEFORE the step's collectives -- one rank overflowing must stop all of them, together
        cap = 1000
        count = 990 + 3 * rank
        for step in range(6):
            count += 4 * rank  # rank 1 gains particles, rank 0 does not
            ok = slabs.agree(count <= cap, dist)
            out.append(ok)
            if not ok:
                break
            t = torch.ones(1)
            dist.all_reduce(t)  # the "step": a collective every rank must enter
        ret[rank] = out
    except BaseException as ex:  # noqa: BLE001
        ret[rank] = "%s: %s" % (type(ex).__name__, ex)
    finally:
        dist.destroy_process_group()


def test_ranks_agree_on_local_failures_over_gloo():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_agree_worker, args=(r, 2, port, ret)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(not p.is_alive() for p in procs)  # nobody is left waiting in a collective
    assert ret[0] == ret[1] == [True, False]     # rank 1 overflows at its second step; rank 0 stops with it


def test_multigrid_layout_is_rank_independent():
    # the bench grids: 8 slabs of 64 layers coarsen to 1 layer per rank; agglomeration starts at the 64^3 level
    levels, agg = slabs.multigrid_layout(512, 512, 512, 8)
    assert levels[0] == (512, 512, 512) and levels[-1] == (8, 8, 8) and len(levels) == 7 and agg == 3
    levels, agg = slabs.multigrid_layout(512, 256, 256, 2)
    assert agg == 2 and levels[agg] == (128, 64, 64)
    # unequal, odd slabs: level 0 is the only distributed level, nothing to agglomerate
    levels, agg = slabs.multigrid_layout(24, 20, 19, 2)
    assert levels == [(24, 20, 19)] and agg is None
    # one GPU: down to 2^3, no agglomeration
    levels, agg = slabs.multigrid_layout(256, 256, 256, 1)
    assert levels[-1] == (2, 2, 2) and agg is None


def test_bench_workload_shapes():
    import bench
    assert bench.workload_dims(256, 1) == (256, 256, 256)
    assert bench.workload_dims(256, 2) == (512, 256, 256)
    assert bench.workload_dims(256, 4) == (512, 512, 256)
    assert bench.workload_dims(256, 8) == (512, 512, 512)       # BASELINE configs[3]
    for w in (1, 2, 4, 8, 3):
        nx, ny, nz = bench.workload_dims(256, w)
        assert nx * ny * nz == w * 256 ** 3                     # weak scaling: 256^3 cells per GPU
    boxes = bench.scene_boxes(512, 512, 512)
    assert boxes[0][1][2] == 512.0 and boxes[0][1][1] < 512.0   # the water spans z, leaves air on top
