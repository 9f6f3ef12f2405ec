"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the plain-C restatement oracle/fluid_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libfluid_oracle.so")

AIR, FLUID, SOLID = 1, 2, 4
PIC, FLIP, APIC = 0, 1, 2
NOT_FLUID = np.uint64(0xFFFFFFFFFFFFFFFF)


class Params(C.Structure):
    _fields_ = [("nx", C.c_uint64), ("ny", C.c_uint64), ("nz", C.c_uint64), ("h", C.c_double),
                ("off", C.c_double * 3), ("g", C.c_double * 3), ("rho", C.c_double), ("skin", C.c_double),
                ("stiffness", C.c_double), ("blend", C.c_double), ("method", C.c_int32),
                ("extrap_iters", C.c_int32)]


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        vp, sz, db = C.c_void_p, C.c_size_t, C.c_double
        pp = C.POINTER(Params)
        L.fo_cell_keys.argtypes = [pp, sz, vp, vp]
        L.fo_hash.restype = sz
        L.fo_hash.argtypes = [pp, sz, vp, vp, vp, vp, vp]
        L.fo_cell_ranges.restype = sz
        L.fo_cell_ranges.argtypes = [pp, sz, vp, vp, vp, vp]
        L.fo_p2g.argtypes = [pp, sz, vp, vp, vp, vp, vp, vp, vp, vp]
        L.fo_gravity.argtypes = [pp, db, vp]
        L.fo_solver_setup.argtypes = [pp, vp, vp, sz, vp, vp, vp, vp]
        L.fo_apply_a.argtypes = [pp, db, sz, vp, vp, vp, vp, vp]
        L.fo_solve.restype = sz
        L.fo_solve.argtypes = [pp, db, sz, vp, vp, vp, vp, db, db, db, sz, vp, vp]
        L.fo_apply_pressure.argtypes = [pp, db, sz, vp, vp, vp, vp, vp]
        L.fo_extrapolate.argtypes = [pp, sz, vp, vp, vp]
        L.fo_g2p.argtypes = [pp, sz, vp, vp, vp, vp, vp]
        L.fo_advect.argtypes = [pp, db, sz, vp, vp]
        L.fo_collide.argtypes = [pp, sz, vp, vp, vp]
        L.fo_correct.argtypes = [pp, db, sz, vp, vp, vp]
        L.fo_cfl.restype = db
        L.fo_cfl.argtypes = [pp, sz, vp]
        L.fo_time_step.restype = sz
        L.fo_time_step.argtypes = [pp, db, sz, vp, vp, vp, vp, vp, vp, vp, db, sz, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    """Stateless helpers around the C restatement; arrays are numpy, contiguous, modified in place where the
    reference modifies in place."""

    def __init__(self, size, h=1.0, offset=(0, 0, 0), gravity=(0, -981.0, 0), method=APIC, blend=1.0,
                 density=1.0, skin=0.1, stiffness=5.0, extrap_iters=1):
        self.L = lib()
        self.P = Params()
        self.P.nx, self.P.ny, self.P.nz = (int(s) for s in size)
        self.P.h = h
        self.P.off[:] = [float(v) for v in offset]
        self.P.g[:] = [float(v) for v in gravity]
        self.P.rho, self.P.skin, self.P.stiffness, self.P.blend = density, skin, stiffness, blend
        self.P.method, self.P.extrap_iters = int(method), int(extrap_iters)
        self.nc = int(self.P.nx * self.P.ny * self.P.nz)

    def cell_keys(self, pos):
        pos = f64(pos)
        key = np.zeros(pos.shape[0], dtype=np.uint64)
        self.L.fo_cell_keys(self.P, pos.shape[0], _p(pos), _p(key))
        return key

    def hash(self, key):
        key = np.ascontiguousarray(key, dtype=np.uint64)
        n = key.shape[0]
        perm = np.zeros(n, dtype=np.uint64)
        begin = np.zeros(self.nc, dtype=np.uint64)
        count = np.zeros(self.nc, dtype=np.uint64)
        fluid = np.zeros(min(n, self.nc), dtype=np.uint64)
        nf = self.L.fo_hash(self.P, n, _p(key), _p(perm), _p(begin), _p(count), _p(fluid))
        return perm, begin, count, fluid[:nf].copy()

    def cell_ranges(self, sorted_key):
        key = np.ascontiguousarray(sorted_key, dtype=np.uint64)
        n = key.shape[0]
        begin = np.zeros(self.nc, dtype=np.uint64)
        count = np.zeros(self.nc, dtype=np.uint64)
        fluid = np.zeros(min(n, self.nc), dtype=np.uint64)
        nf = self.L.fo_cell_ranges(self.P, n, _p(key), _p(begin), _p(count), _p(fluid))
        return begin, count, fluid[:nf].copy()

    def p2g(self, pos, vel, c, begin, count, gvel, types, old_gvel=None):
        pos, vel, c = f64(pos), f64(vel), f64(c)
        if self.P.method == FLIP and old_gvel is None:
            raise ValueError("FLIP needs old_gvel")
        self.L.fo_p2g(self.P, pos.shape[0], _p(pos), _p(vel), _p(c), _p(begin), _p(count), _p(gvel), _p(types),
                      _p(old_gvel))

    def gravity(self, dt, gvel):
        self.L.fo_gravity(self.P, dt, _p(gvel))

    def solver_setup(self, gvel, types, fluid_cells):
        nf = fluid_cells.shape[0]
        index_map = np.zeros(self.nc, dtype=np.uint64)
        flags = np.zeros(nf, dtype=np.uint8)
        b = np.zeros(nf, dtype=np.float64)
        self.L.fo_solver_setup(self.P, _p(gvel), _p(types), nf, _p(fluid_cells), _p(index_map), _p(flags), _p(b))
        return index_map, flags, b

    def apply_a(self, a_scale, fluid_cells, index_map, flags, v):
        v = f64(v)
        out = np.zeros_like(v)
        self.L.fo_apply_a(self.P, a_scale, v.shape[0], _p(fluid_cells), _p(index_map), _p(flags), _p(v), _p(out))
        return out

    def solve(self, dt, fluid_cells, index_map, flags, b, tau=0.97, sigma=0.25, tolerance=1e-6,
              max_iterations=200):
        nf = fluid_cells.shape[0]
        p = np.zeros(nf, dtype=np.float64)
        res = C.c_double(0.0)
        it = self.L.fo_solve(self.P, dt, nf, _p(fluid_cells), _p(index_map), _p(flags), _p(b), tau, sigma,
                             tolerance, max_iterations, _p(p), C.byref(res))
        return p, res.value, it

    def apply_pressure(self, dt, fluid_cells, index_map, p, gvel, types):
        self.L.fo_apply_pressure(self.P, dt, fluid_cells.shape[0], _p(fluid_cells), _p(index_map), _p(f64(p)),
                                 _p(gvel), _p(types))

    def extrapolate(self, fluid_cells, gvel, types):
        self.L.fo_extrapolate(self.P, fluid_cells.shape[0], _p(fluid_cells), _p(gvel), _p(types))

    def g2p(self, pos, vel, c, gvel, old_gvel=None):
        self.L.fo_g2p(self.P, pos.shape[0], _p(pos), _p(vel), _p(c), _p(gvel), _p(old_gvel))

    def advect(self, dt, pos, vel):
        self.L.fo_advect(self.P, dt, pos.shape[0], _p(pos), _p(vel))

    def collide(self, pos, old_pos, types):
        self.L.fo_collide(self.P, pos.shape[0], _p(pos), _p(old_pos), _p(types))

    def correct(self, dt, pos, begin, count):
        self.L.fo_correct(self.P, dt, pos.shape[0], _p(pos), _p(begin), _p(count))

    def cfl(self, vel):
        return self.L.fo_cfl(self.P, vel.shape[0], _p(vel))

    def time_step(self, dt, pos, vel, c, old_pos, gvel, types, old_gvel, tolerance=1e-6, max_iterations=200,
                  phases=None):
        res = C.c_double(0.0)
        it = self.L.fo_time_step(self.P, dt, pos.shape[0], _p(pos), _p(vel), _p(c), _p(old_pos), _p(gvel),
                                 _p(types), _p(old_gvel), tolerance, max_iterations, C.byref(res), _p(phases))
        return it, res.value
