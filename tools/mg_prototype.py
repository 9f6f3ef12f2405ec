"""CPU prototype used to choose the multigrid preconditioner that replaces the reference's sequential MIC(0).
Builds the reference's pressure matrix (diag = #non-solid neighbours, -1 to fluid neighbours) on a synthetic
free-surface scene with an obstacle and counts PCG iterations for several V-cycle variants.
Usage: python tools/mg_prototype.py [n] [scenes]            V-cycle variants
       python tools/mg_prototype.py [n] [scenes] --experiments   storage precision / per-level sweeps"""
import sys
import time

import numpy as np
import scipy.sparse as sp

AIR, FLUID, SOLID = 1, 2, 4


def scene(n, kind="wave"):
    z, y, x = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    t = np.full((n, n, n), AIR, dtype=np.uint8)
    if kind == "wave":
        surf = 0.55 * n + 0.18 * n * np.sin(2 * np.pi * x / n) * np.cos(2 * np.pi * z / n)
        t[y < surf] = FLUID
        t[(abs(x - 0.5 * n) < 0.1 * n) & (abs(z - 0.5 * n) < 0.1 * n) & (y < 0.25 * n)] = SOLID
    elif kind == "dam":
        t[(x < 0.2 * n)] = FLUID
    elif kind == "full":
        t[y < n - 1] = FLUID
    elif kind == "splash":  # thin sheets + droplets
        rng = np.random.default_rng(0)
        t[y < 0.2 * n] = FLUID
        t[(rng.random((n, n, n)) < 0.05) & (y < 0.8 * n)] = FLUID
        t[(abs(x - 0.3 * n) < 2) & (y < 0.7 * n)] = FLUID
    return t


def build_matrix(t):
    n = t.shape[0]
    tp = np.pad(t, 1, constant_values=SOLID)
    fluid = t == FLUID
    idx = -np.ones(t.shape, dtype=np.int64)
    idx[fluid] = np.arange(fluid.sum())
    diag = np.zeros(t.shape)
    rows, cols, vals = [], [], []
    for ax in range(3):
        for s in (-1, 1):
            sl = [slice(1, -1)] * 3
            sl[ax] = slice(1 + s, tp.shape[ax] - 1 + s)
            nb = tp[tuple(sl)]
            diag += nb != SOLID
            m = fluid & (nb == FLUID)
            nbidx = np.roll(idx, -s, axis=ax)
            rows.append(idx[m]); cols.append(nbidx[m]); vals.append(-np.ones(m.sum()))
    nf = fluid.sum()
    rows.append(np.arange(nf)); cols.append(np.arange(nf)); vals.append(diag[fluid])
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nf, nf))
    coords = np.stack(np.nonzero(fluid), axis=1)  # z, y, x
    return A, coords


def aggregate(coords):
    cc = coords // 2
    key = (cc[:, 0].astype(np.int64) << 40) | (cc[:, 1].astype(np.int64) << 20) | cc[:, 2]
    uniq, inv = np.unique(key, return_inverse=True)
    P = sp.csr_matrix((np.ones(len(inv)), (np.arange(len(inv)), inv)), shape=(len(inv), len(uniq)))
    ccoords = np.stack([uniq >> 40, (uniq >> 20) & 0xFFFFF, uniq & 0xFFFFF], axis=1)
    return P, ccoords


class Level:
    pass


def setup(A, coords, min_size=50, max_levels=12):
    levels = []
    while True:
        L = Level()
        L.A = A.tocsr()
        L.D = A.diagonal()
        L.color = (coords.sum(axis=1) & 1).astype(bool)
        levels.append(L)
        if A.shape[0] <= min_size or len(levels) >= max_levels:
            break
        P, ccoords = aggregate(coords)
        L.P = P
        A = (P.T @ A @ P).tocsr()
        coords = ccoords
    return levels


def rbgs(L, x, b, order, sweeps):
    for _ in range(sweeps):
        for col in order:
            m = L.color == col
            r = b - L.A @ x
            x[m] += r[m] / L.D[m]
    return x


def jacobi(L, x, b, sweeps, w=0.8):
    for _ in range(sweeps):
        x = x + w * (b - L.A @ x) / L.D
    return x


def vcycle(levels, l, b, cfg):
    L = levels[l]
    x = np.zeros_like(b)
    if l == len(levels) - 1:
        if L.A.shape[0] <= 2000:
            return np.linalg.solve(L.A.toarray(), b) if cfg.get("exact_coarse", False) else rbgs(L, x, b, (False, True, True, False), cfg.get("coarse_sweeps", 8))
        return rbgs(L, x, b, (False, True, True, False), cfg.get("coarse_sweeps", 8))
    if cfg["smoother"] == "rbgs":
        x = rbgs(L, x, b, (False, True), cfg["pre"])
    else:
        x = jacobi(L, x, b, cfg["pre"])
    r = b - L.A @ x
    rc = L.P.T @ r
    ec = vcycle(levels, l + 1, rc, cfg)
    if cfg.get("wcycle", False) and l + 1 < len(levels) - 1 and l >= cfg.get("wfrom", 0):
        r2 = rc - levels[l + 1].A @ ec
        ec = ec + vcycle(levels, l + 1, r2, cfg)
    x = x + cfg.get("omega", 1.0) * (L.P @ ec)
    if cfg["smoother"] == "rbgs":
        x = rbgs(L, x, b, (True, False), cfg["post"])
    else:
        x = jacobi(L, x, b, cfg["post"])
    return x


def pcg(A, b, M, tol_abs, maxit=2000):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    s = z.copy()
    sigma = z @ r
    for it in range(1, maxit + 1):
        q = A @ s
        alpha = sigma / (q @ s)
        x += alpha * s
        r -= alpha * q
        if np.abs(r).max() < tol_abs:
            return x, it
        z = M(r)
        sn = z @ r
        s = z + (sn / sigma) * s
        sigma = sn
    return x, maxit


# ---- experiments behind DESIGN.md section 3.2 (python tools/mg_prototype.py 64 wave,full --experiments) -------------
def _quant(dtype):
    return lambda v: v.astype(dtype).astype(np.float64)


def _rbgs_q(L, x, b, order, sweeps, q):
    for _ in range(sweeps):
        for col in order:
            msk = L.color == col
            r = b - L.A @ x
            x[msk] += r[msk] / L.D[msk]
            x = q(x)
    return x


def vcycle_storage(levels, l, b, q0, pp=((2, 2),), coarse_sweeps=8, omega=1.8):
    """V-cycle whose level-0 vectors are stored through q0 (fp16 / fp32 rounding after every half-sweep), the other
    levels in fp32; pp[l] = (pre, post) sweeps of level l (last entry repeats)."""
    L = levels[l]
    q = q0 if l == 0 else _quant(np.float32)
    x = np.zeros_like(b)
    if l == len(levels) - 1:
        return rbgs(L, x, b, (False, True, True, False), coarse_sweeps)
    pre, post = pp[min(l, len(pp) - 1)]
    x = _rbgs_q(L, x, b, (False, True), pre, q)
    ec = vcycle_storage(levels, l + 1, L.P.T @ (b - L.A @ x), q0, pp, coarse_sweeps, omega)
    x = q(x + omega * (L.P @ ec))
    return _rbgs_q(L, x, b, (True, False), post, q)


def precond_scaled(levels, r, q0, **kw):
    """the level-0 right-hand side is r / max|r| so that it fits a 16-bit float at every stage of the solve"""
    s = np.abs(r).max()
    return r if s == 0 else vcycle_storage(levels, 0, q0(r / s), q0, **kw) * s


def experiments(n, kinds):
    for kind in kinds:
        A, coords = build_matrix(scene(n, kind))
        b = np.random.default_rng(1).uniform(-30, 30, A.shape[0])
        levels = setup(A, coords)
        print("== %s n=%d unknowns=%d" % (kind, n, A.shape[0]))
        for name, q in (("level-0 vectors fp64", lambda v: v), ("level-0 vectors fp32 (production)", _quant(np.float32)),
                        ("level-0 vectors fp16, rhs scaled by max|r|", _quant(np.float16))):
            _, it = pcg(A, b, lambda r: precond_scaled(levels, r, q), 1e-6)
            print("   %-50s iters %3d" % (name, it))
        for name, pp in (("V(2,2) on every level (production)", ((2, 2),)), ("level 0 (1,1), coarser (2,2)", ((1, 1), (2, 2))),
                         ("level 0 (2,2), coarser (3,3)", ((2, 2), (3, 3))), ("level 0 (2,2), coarser (4,4)", ((2, 2), (4, 4)))):
            _, it = pcg(A, b, lambda r: precond_scaled(levels, r, _quant(np.float32), pp=pp), 1e-6)
            print("   %-50s iters %3d" % (name, it))
        for cs in (8, 4, 2):
            _, it = pcg(A, b, lambda r: precond_scaled(levels, r, _quant(np.float32), coarse_sweeps=cs), 1e-6)
            print("   %-50s iters %3d" % ("coarsest level: %d symmetric sweeps" % cs, it))


if __name__ == "__main__":
    if "--experiments" in sys.argv:
        sys.argv.remove("--experiments")
        experiments(int(sys.argv[1]) if len(sys.argv) > 1 else 48,
                    sys.argv[2].split(",") if len(sys.argv) > 2 else ["wave", "full"])
        sys.exit(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["wave", "dam", "full", "splash"]
    for kind in kinds:
        t = scene(n, kind)
        A, coords = build_matrix(t)
        rng = np.random.default_rng(1)
        b = rng.uniform(-30, 30, A.shape[0])
        print("== %s n=%d unknowns=%d" % (kind, n, A.shape[0]))
        t0 = time.time()
        _, it = pcg(A, b, lambda r: r / A.diagonal(), 1e-6)
        print("   jacobi-PCG                 iters %4d  (%.1fs)" % (it, time.time() - t0))
        levels = setup(A, coords)
        print("   levels:", [L.A.shape[0] for L in levels])
        for cfg in (
            dict(smoother="rbgs", pre=1, post=1),
            dict(smoother="rbgs", pre=2, post=2),
            dict(smoother="rbgs", pre=1, post=1, omega=1.5),
            dict(smoother="rbgs", pre=2, post=2, omega=1.5),
            dict(smoother="rbgs", pre=2, post=2, omega=1.8),
            dict(smoother="rbgs", pre=1, post=1, wcycle=True),
            dict(smoother="rbgs", pre=2, post=2, wcycle=True, wfrom=1),
            dict(smoother="jacobi", pre=2, post=2, omega=1.5),
        ):
            t0 = time.time()
            _, it = pcg(A, b, lambda r: vcycle(levels, 0, r, cfg), 1e-6)
            print("   %-60s iters %4d  (%.1fs)" % (cfg, it, time.time() - t0))
