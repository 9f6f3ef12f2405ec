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
