"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libfluid_ref.so.

The shared object is the UNMODIFIED reference hot path (lukedan/libfluid src/simulation.cpp,
src/mac_grid.cpp, src/pressure_solver.cpp) compiled by oracle/Makefile together with the phase-replay
driver oracle/ref_driver.cpp.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; the product path (libfluid_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libfluid_ref.so")

# AoS layouts of the reference: include/fluid/simulation.h:24-34 (152 B), include/fluid/mac_grid.h:15-27 (32 B)
PARTICLE_DTYPE = np.dtype([
    ("position", "<f8", 3), ("velocity", "<f8", 3), ("cx", "<f8", 3), ("cy", "<f8", 3), ("cz", "<f8", 3),
    ("old_position", "<f8", 3), ("raw_cell_index", "<u8"),
])
CELL_DTYPE = np.dtype([("vel", "<f8", 3), ("type", "u1"), ("pad", "u1", 7)])
assert PARTICLE_DTYPE.itemsize == 152 and CELL_DTYPE.itemsize == 32

AIR, FLUID, SOLID = 1, 2, 4
PIC, FLIP, APIC = 0, 1, 2


def available():
    return os.path.exists(REF_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_SO)
        vp, sz, db = C.c_void_p, C.c_size_t, C.c_double
        L.ref_create.restype = vp
        L.ref_create.argtypes = [sz, sz, sz, db]
        L.ref_destroy.argtypes = [vp]
        L.ref_set_params.argtypes = [vp, vp, vp, C.c_int, db, db, db, db, sz, db]
        L.ref_seed_box.argtypes = [vp, vp, vp, vp, sz]
        L.ref_seed_sphere.argtypes = [vp, vp, db, vp, sz]
        L.ref_add_source.argtypes = [vp, vp, sz, vp, sz, C.c_int]
        L.ref_num_particles.restype = sz
        L.ref_num_particles.argtypes = [vp]
        for name in ("ref_get_particles", "ref_get_cells", "ref_get_old_cells", "ref_set_cells",
                     "ref_set_old_cells", "ref_get_space_hash", "ref_get_fluid_cells", "ref_get_pressure"):
            getattr(L, name).argtypes = [vp, vp]
        L.ref_set_particles.argtypes = [vp, vp, sz]
        L.ref_set_pressure.argtypes = [vp, vp, sz]
        L.ref_num_fluid_cells.restype = sz
        L.ref_num_fluid_cells.argtypes = [vp]
        for name in ("ref_update_and_hash", "ref_hash", "ref_reset_space_hash", "ref_collide",
                     "ref_save_old_positions", "ref_update_sources", "ref_p2g", "ref_extrapolate", "ref_g2p",
                     "ref_time_step_default"):
            getattr(L, name).argtypes = [vp]
        for name in ("ref_advect", "ref_gravity", "ref_apply_pressure", "ref_correct", "ref_time_step",
                     "ref_update"):
            getattr(L, name).argtypes = [vp, db]
        L.ref_solver_setup.argtypes = [vp, db, sz]
        L.ref_solve.argtypes = [vp, db, vp, vp]
        L.ref_solver_rhs.argtypes = [vp, db, vp, vp]
        L.ref_solver_apply_a.argtypes = [vp, vp, vp]
        L.ref_cfl.restype = db
        L.ref_cfl.argtypes = [vp]
        if hasattr(L, "ref_mesher_sample"):
            L.ref_mesher_sample.argtypes = [vp, vp, db, db, sz, vp, sz, db, vp]
            L.ref_voxelize.argtypes = [vp, sz, vp, sz, db, vp, vp, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _v3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(3))


class RefSim:
    """The reference fluid::simulation, driven headlessly phase by phase."""

    def __init__(self, size, h=1.0, offset=(0, 0, 0), gravity=(0, -981.0, 0), method=APIC, blend=1.0,
                 density=1.0, skin=0.1, stiffness=5.0, extrap_iters=1, cfl_number=3.0):
        self.L = lib()
        self.size = tuple(int(s) for s in size)
        self.h = float(h)
        self.ptr = self.L.ref_create(*self.size, self.h)
        self.set_params(offset, gravity, method, blend, density, skin, stiffness, extrap_iters, cfl_number)

    def __del__(self):
        if getattr(self, "ptr", None):
            self.L.ref_destroy(self.ptr)
            self.ptr = None

    def set_params(self, offset, gravity, method, blend=1.0, density=1.0, skin=0.1, stiffness=5.0,
                   extrap_iters=1, cfl_number=3.0):
        self.offset, self.gravity, self.method = _v3(offset), _v3(gravity), int(method)
        self.blend, self.density, self.skin, self.stiffness = blend, density, skin, stiffness
        self.extrap_iters, self.cfl_number = extrap_iters, cfl_number
        self.L.ref_set_params(self.ptr, _p(self.offset), _p(self.gravity), self.method, blend, density, skin,
                              stiffness, extrap_iters, cfl_number)

    @property
    def ncells(self):
        return self.size[0] * self.size[1] * self.size[2]

    # ---- state ----
    def seed_box(self, start, size, vel=(0, 0, 0), dens=2):
        self.L.ref_seed_box(self.ptr, _p(_v3(start)), _p(_v3(size)), _p(_v3(vel)), dens)

    def seed_sphere(self, center, radius, vel=(0, 0, 0), dens=2):
        self.L.ref_seed_sphere(self.ptr, _p(_v3(center)), radius, _p(_v3(vel)), dens)

    def add_source(self, cells_xyz, vel, dens=2, coerce=False):
        a = np.ascontiguousarray(np.asarray(cells_xyz, dtype=np.uint64).reshape(-1, 3))
        self.L.ref_add_source(self.ptr, _p(a), a.shape[0], _p(_v3(vel)), dens, int(coerce))

    def num_particles(self):
        return self.L.ref_num_particles(self.ptr)

    def particles(self):
        out = np.empty(self.num_particles(), dtype=PARTICLE_DTYPE)
        self.L.ref_get_particles(self.ptr, _p(out))
        return out

    def set_particles(self, arr):
        arr = np.ascontiguousarray(arr, dtype=PARTICLE_DTYPE)
        self.L.ref_set_particles(self.ptr, _p(arr), arr.shape[0])

    def cells(self):
        out = np.zeros(self.ncells, dtype=CELL_DTYPE)
        self.L.ref_get_cells(self.ptr, _p(out))
        return out

    def set_cells(self, arr):
        arr = np.ascontiguousarray(arr, dtype=CELL_DTYPE)
        assert arr.shape[0] == self.ncells
        self.L.ref_set_cells(self.ptr, _p(arr))

    def old_cells(self):
        out = np.zeros(self.ncells, dtype=CELL_DTYPE)
        self.L.ref_get_old_cells(self.ptr, _p(out))
        return out

    def set_old_cells(self, arr):
        arr = np.ascontiguousarray(arr, dtype=CELL_DTYPE)
        self.L.ref_set_old_cells(self.ptr, _p(arr))

    def set_solid(self, mask_zyx):
        """mask indexed [z, y, x] (raw order x fastest)."""
        c = self.cells()
        c["type"][np.asarray(mask_zyx).reshape(-1)] = SOLID
        self.set_cells(c)

    def space_hash(self):
        out = np.zeros((self.ncells, 2), dtype=np.uint64)
        self.L.ref_get_space_hash(self.ptr, _p(out))
        return out[:, 0].copy(), out[:, 1].copy()

    def fluid_cells(self):
        out = np.zeros(self.L.ref_num_fluid_cells(self.ptr), dtype=np.uint64)
        self.L.ref_get_fluid_cells(self.ptr, _p(out))
        return out

    # ---- phases (reference src/simulation.cpp:43-125) ----
    def update_and_hash(self):
        self.L.ref_update_and_hash(self.ptr)

    def hash(self):
        self.L.ref_hash(self.ptr)

    def reset_space_hash(self):
        self.L.ref_reset_space_hash(self.ptr)

    def advect(self, dt):
        self.L.ref_advect(self.ptr, dt)

    def collide(self):
        self.L.ref_collide(self.ptr)

    def save_old_positions(self):
        self.L.ref_save_old_positions(self.ptr)

    def update_sources(self):
        self.L.ref_update_sources(self.ptr)

    def p2g(self):
        self.L.ref_p2g(self.ptr)

    def add_gravity(self, dt):
        self.L.ref_gravity(self.ptr, dt)

    def solver_setup(self, tolerance=-1.0, max_iterations=0):
        self.L.ref_solver_setup(self.ptr, tolerance, max_iterations)

    def solver_rhs(self, dt):
        nf = self.L.ref_num_fluid_cells(self.ptr)
        b = np.zeros(nf, dtype=np.float64)
        fl = np.zeros(nf, dtype=np.uint8)
        self.L.ref_solver_rhs(self.ptr, dt, _p(b), _p(fl))
        return b, fl

    def solver_apply_a(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.zeros_like(v)
        self.L.ref_solver_apply_a(self.ptr, _p(v), _p(out))
        return out

    def solve(self, dt):
        res = C.c_double(0.0)
        it = C.c_size_t(0)
        self.L.ref_solve(self.ptr, dt, C.byref(res), C.byref(it))
        p = np.zeros(self.L.ref_num_fluid_cells(self.ptr), dtype=np.float64)
        self.L.ref_get_pressure(self.ptr, _p(p))
        return p, res.value, it.value

    def set_pressure(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        self.L.ref_set_pressure(self.ptr, _p(p), p.shape[0])

    def apply_pressure(self, dt):
        self.L.ref_apply_pressure(self.ptr, dt)

    def correct(self, dt):
        self.L.ref_correct(self.ptr, dt)

    def extrapolate(self):
        self.L.ref_extrapolate(self.ptr)

    def g2p(self):
        self.L.ref_g2p(self.ptr)

    def cfl(self):
        return self.L.ref_cfl(self.ptr)

    def time_step(self, dt=None):
        if dt is None:
            self.L.ref_time_step_default(self.ptr)
        else:
            self.L.ref_time_step(self.ptr, dt)

    def update(self, dt):
        self.L.ref_update(self.ptr, dt)

    def replay_time_step(self, dt, hook=None, tolerance=-1.0, max_iterations=0):
        """time_step(dt) replayed phase by phase (bit-identical to the stock call; tests check it).

        hook(name, self) is called after every phase so that callers can snapshot state."""
        def h(name):
            if hook:
                hook(name, self)
        self.update_and_hash(); h("hash0")
        self.advect(dt); h("advect")
        self.collide(); self.save_old_positions(); h("collide1")
        self.update_and_hash()
        self.update_sources()
        self.hash(); h("hash")
        self.p2g(); h("p2g")
        self.add_gravity(dt); h("gravity")
        self.solver_setup(tolerance, max_iterations)
        out = self.solve(dt); h("solve")
        self.apply_pressure(dt); h("apply_pressure")
        self.correct(dt); h("correct")
        self.collide(); self.save_old_positions(); h("collide2")
        self.extrapolate(); h("extrapolate")
        self.g2p(); h("g2p")
        return out


def mesher_sample(size, offset, cell_size, extent, cell_radius, xyz, r):
    """mesher::_sample_surface_function of the reference: (sz + 1, sy + 1, sx + 1) array of the implicit function"""
    L = lib()
    size = np.ascontiguousarray(size, dtype=np.uint64)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
    out = np.zeros((int(size[2]) + 1, int(size[1]) + 1, int(size[0]) + 1), dtype=np.float64)
    L.ref_mesher_sample(_p(size), _p(_v3(offset)), float(cell_size), float(extent), int(cell_radius), _p(xyz),
                        xyz.shape[0], float(r), _p(out))
    return out


def voxelize(positions, indices, cell_size, ref_offset, ref_size):
    """voxelizer + obstacle of the reference: (grid_min (3,), voxels (vz, vy, vx) u8, obstacle cells (n, 3))"""
    L = lib()
    pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
    idx = np.ascontiguousarray(indices, dtype=np.uint64).ravel()
    rsz = np.ascontiguousarray(ref_size, dtype=np.uint64)
    gmin, vsz, nc = np.zeros(3, dtype=np.int64), np.zeros(3, dtype=np.uint64), C.c_size_t()
    L.ref_voxelize(_p(pos), pos.shape[0], _p(idx), idx.shape[0], float(cell_size), _p(_v3(ref_offset)), _p(rsz), _p(gmin),
                   _p(vsz), None, C.byref(nc), None)
    vox = np.zeros((int(vsz[2]), int(vsz[1]), int(vsz[0])), dtype=np.uint8)
    undefined = nc.value == 2 ** 64 - 1  # the reference's obstacle would read out of bounds (see ref_driver.cpp)
    cells = np.zeros((1 if undefined else max(nc.value, 1), 3), dtype=np.uint64)
    L.ref_voxelize(_p(pos), pos.shape[0], _p(idx), idx.shape[0], float(cell_size), _p(_v3(ref_offset)), _p(rsz), _p(gmin),
                   _p(vsz), _p(vox), C.byref(nc), _p(cells))
    return gmin, vox, (None if undefined else cells[:nc.value])
