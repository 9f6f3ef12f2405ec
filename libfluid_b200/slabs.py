"""Host-side mirror of the multi-GPU decomposition rules of the CUDA library (lfk_create / exchange.cu), used by
bench.py, the multi-GPU check and the CPU (gloo) tests of the rank protocol.

The grid is cut into contiguous z-slabs (z is the slowest raw index, include/fluid/data_structures/grid.h:24-31 of
the reference, so a slab is a contiguous raw range).  Rank r owns z in [z0, z0 + nzl); the remainder of nz / nranks
goes to the first ranks; slabs must be >= MIN_SLAB cells thick because a particle travels at most cfl_number = 3
cells per step (simulation.h:183, src/simulation.cpp:33): an immigrant then cannot reach the far boundary layer of
its new slab in the step it arrives in, so ONE neighbour exchange per step settles ownership and ghost copies.
"""
import numpy as np

MIN_SLAB = 4


def slab_range(nz, nranks, rank):
    """(z0, nzl) exactly as lfk_create computes them."""
    if nranks < 1 or not 0 <= rank < nranks or (nranks > 1 and nz < MIN_SLAB * nranks):
        raise ValueError("bad rank layout (slabs must be >= %d cells thick)" % MIN_SLAB)
    base, rem = divmod(nz, nranks)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def owner_of_z(nz, nranks, zc):
    """rank owning z cell(s) zc"""
    zc = np.asarray(zc)
    base, rem = divmod(nz, nranks)
    split = rem * (base + 1)
    return np.where(zc < split, zc // (base + 1), rem + (zc - split) // max(base, 1)).astype(np.int64)


def z_cell(pz, nz, h=1.0, off=0.0):
    """K1 along z (src/simulation.cpp:251-261): trunc(max((p - off) / h, 0)) clamped to nz - 1"""
    g = np.maximum((np.asarray(pz, dtype=np.float64) - off) / h, 0.0)
    return np.minimum(g.astype(np.uint64), np.uint64(nz - 1)).astype(np.int64)


def classify(zc, z0, nzl, has_up, has_dn):
    """The exchange rule of exchange.cu (k_xch_*): boolean masks (send_up, send_down, dead) for own particles."""
    zc = np.asarray(zc)
    top = z0 + nzl
    up = (zc >= top - 1) if has_up else np.zeros(zc.shape, bool)
    dn = (zc <= z0) if has_dn else np.zeros(zc.shape, bool)
    dead = (zc > top) | (zc < z0 - 1)
    return up, dn, dead


def split_after_sort(zc, z0, nzl):
    """What the cell sort makes of the merged set: masks (ghost_low, own, ghost_high)."""
    zc = np.asarray(zc)
    return zc == z0 - 1, (zc >= z0) & (zc < z0 + nzl), zc == z0 + nzl


def agree(ok, dist=None, device=None):
    """True iff EVERY rank passed True.  A rank that fails a local check (buffer too small, allocation failed) must not
    raise on its own while its peers walk into the step's collectives (the round-1 hang of bench.py --gpus 2): every
    rank learns the outcome first and all of them leave together."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return bool(ok)
    import torch
    t = torch.tensor([0.0 if ok else 1.0], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) == 0.0


def multigrid_layout(nx, ny, nz, nranks, agg_max_cells=600000, tail_cells=4096):
    """Host mirror of mg_alloc / mg_agg_alloc (libfluid_b200/csrc/mg.cu): the distributed multigrid levels
    [(nx, ny, nz_global)], and the index of the first level that is agglomerated onto every rank (None: none).  It is a
    pure function of the whole grid and the rank count -- every rank must build the same hierarchy, or the exchanges
    inside the V-cycle would not pair up."""
    layout = [slab_range(nz, nranks, r) for r in range(nranks)]
    z0 = [a for a, _ in layout]
    nzl = [b for _, b in layout]
    levels = []
    gx, gy, gz = nx, ny, nz
    for _ in range(16):
        levels.append((gx, gy, gz))
        aligned = True
        if nranks > 1:
            aligned = gz % 2 == 0 and all(a % 2 == 0 and b % 2 == 0 and b >= 2 for a, b in zip(z0, nzl))
        lx, ly, lz = (gx, gy, max(nzl)) if nranks > 1 else (gx, gy, gz)
        if not aligned or (lx <= 2 and ly <= 2 and lz <= 2):
            break
        gx, gy, gz = (gx + 1) // 2, (gy + 1) // 2, (gz + 1) // 2
        z0 = [a // 2 for a in z0]
        nzl = [(b + 1) // 2 for b in nzl]
    agg = None
    if nranks > 1:
        for l in range(1, len(levels)):
            if levels[l][0] * levels[l][1] * (nz >> l) <= agg_max_cells:
                agg = l
                break
    return levels, agg
