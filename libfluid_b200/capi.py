"""ctypes binding of the lfk C ABI (include/lfk.h) -- the host-side handle used by tests, bench.py and the Python
mirror of the reference interface (libfluid_b200.simulation).

There is no CPU fallback: if the CUDA library is missing or no device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_lib", "liblfk.so")

AIR, FLUID, SOLID = 1, 2, 4
PIC, FLIP, APIC = 0, 1, 2
PRECOND_JACOBI, PRECOND_MULTIGRID = 0, 1

# host layouts of the reference (include/fluid/simulation.h:24-34, include/fluid/mac_grid.h:15-27)
PARTICLE_DTYPE = np.dtype([
    ("position", "<f8", 3), ("velocity", "<f8", 3), ("cx", "<f8", 3), ("cy", "<f8", 3), ("cz", "<f8", 3),
    ("old_position", "<f8", 3), ("raw_cell_index", "<u8"),
])
CELL_DTYPE = np.dtype([("vel", "<f8", 3), ("type", "u1"), ("pad", "u1", 7)])

PHASES = ("advect_collide", "sort", "p2g", "solve_setup", "pcg", "apply_pressure", "correct_collide",
          "extrapolate", "g2p", "cfl", "transfer", "exchange")


class Params(C.Structure):
    _fields_ = [("grid_offset", C.c_double * 3), ("cell_size", C.c_double), ("density", C.c_double),
                ("gravity", C.c_double * 3), ("boundary_skin_width", C.c_double),
                ("correction_stiffness", C.c_double), ("blending_factor", C.c_double), ("cfl_number", C.c_double),
                ("tolerance", C.c_double), ("method", C.c_int32), ("extrapolation_iterations", C.c_int32),
                ("max_iterations", C.c_int32), ("preconditioner", C.c_int32)]


class Source(C.Structure):
    _fields_ = [("cells", C.c_void_p), ("num_cells", C.c_uint64), ("velocity", C.c_double * 3),
                ("target_density_cubic_root", C.c_uint32), ("active", C.c_int32), ("coerce_velocity", C.c_int32)]


class Mesher(C.Structure):
    _fields_ = [("grid_offset", C.c_double * 3), ("cell_size", C.c_double), ("particle_extent", C.c_double),
                ("cell_radius", C.c_uint64), ("size", C.c_uint64 * 3)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("pcg_iterations", C.c_uint64), ("pcg_residual", C.c_double),
                ("phase_ms", C.c_double * 16), ("num_particles", C.c_uint64), ("num_fluid_cells", C.c_uint64),
                ("exchanged_particles", C.c_uint64)]


class LfkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (code %d)" % (msg, code))
        self.code = code


# every symbol include/lfk.h declares (tests check that the library exports all of them)
SYMBOLS = (
    "lfk_abi_version", "lfk_nccl_unique_id", "lfk_create", "lfk_destroy", "lfk_last_error", "lfk_set_params",
    "lfk_get_params", "lfk_sync", "lfk_slab", "lfk_upload_particles", "lfk_num_particles",
    "lfk_download_particles", "lfk_download_positions", "lfk_download_positions_async", "lfk_wait_transfers",
    "lfk_host_alloc", "lfk_host_free", "lfk_checkpoint_save", "lfk_checkpoint_load", "lfk_upload_cells", "lfk_download_cells",
    "lfk_upload_cells_slab", "lfk_download_cells_slab", "lfk_upload_old_cells", "lfk_download_old_cells", "lfk_download_table", "lfk_num_fluid_cells",
    "lfk_download_fluid_cells", "lfk_hash", "lfk_advect", "lfk_collide", "lfk_p2g", "lfk_gravity",
    "lfk_pressure_solve", "lfk_download_rhs", "lfk_download_pressure", "lfk_upload_pressure", "lfk_apply_a",
    "lfk_apply_pressure", "lfk_correct", "lfk_extrapolate", "lfk_g2p", "lfk_cfl", "lfk_set_sources", "lfk_set_rng_seed",
    "lfk_coerce_sources", "lfk_update_sources", "lfk_mesher_sample", "lfk_voxelize_mesh", "lfk_voxels_download",
    "lfk_obstacle_cells", "lfk_time_step",
    "lfk_time_step_cfl", "lfk_update", "lfk_seed_box_device", "lfk_synthetic_projection_device", "lfk_set_timing",
    "lfk_set_tuning", "lfk_get_stats", "lfk_reset_stats",
)

_lib = None


def load_library():
    """dlopens liblfk.so (built in-tree by libfluid_b200.build).  Raises if it is missing: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LfkError(-1, "liblfk.so is not built (run `python -m libfluid_b200.build`); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u64, db, ci = C.c_void_p, C.c_uint64, C.c_double, C.c_int
    L.lfk_last_error.restype = C.c_char_p
    L.lfk_last_error.argtypes = [vp]
    L.lfk_nccl_unique_id.argtypes = [vp]
    L.lfk_create.argtypes = [C.POINTER(vp), u64, u64, u64, ci, vp, ci, ci, vp]
    L.lfk_destroy.argtypes = [vp]
    L.lfk_set_params.argtypes = [vp, C.POINTER(Params)]
    L.lfk_get_params.argtypes = [vp, C.POINTER(Params)]
    L.lfk_sync.argtypes = [vp]
    L.lfk_slab.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.lfk_upload_particles.argtypes = [vp, vp, u64]
    L.lfk_num_particles.argtypes = [vp, C.POINTER(u64)]
    L.lfk_download_particles.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.lfk_download_positions.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.lfk_download_positions_async.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.lfk_wait_transfers.argtypes = [vp]
    L.lfk_host_alloc.argtypes = [C.POINTER(vp), u64]
    L.lfk_host_free.argtypes = [vp]
    L.lfk_checkpoint_save.argtypes = [vp, C.c_char_p]
    L.lfk_checkpoint_load.argtypes = [vp, C.c_char_p]
    for n in ("lfk_upload_cells", "lfk_download_cells", "lfk_upload_old_cells", "lfk_download_old_cells",
              "lfk_upload_cells_slab", "lfk_download_cells_slab"):
        getattr(L, n).argtypes = [vp, vp]
    L.lfk_download_table.argtypes = [vp, vp, vp]
    L.lfk_num_fluid_cells.argtypes = [vp, C.POINTER(u64)]
    L.lfk_download_fluid_cells.argtypes = [vp, vp, u64]
    for n in ("lfk_hash", "lfk_collide", "lfk_p2g", "lfk_extrapolate", "lfk_g2p", "lfk_reset_stats"):
        getattr(L, n).argtypes = [vp]
    for n in ("lfk_advect", "lfk_gravity", "lfk_apply_pressure", "lfk_correct", "lfk_time_step"):
        getattr(L, n).argtypes = [vp, db]
    L.lfk_pressure_solve.argtypes = [vp, db, C.POINTER(db), C.POINTER(u64)]
    L.lfk_download_rhs.argtypes = [vp, db, vp, vp, u64]
    L.lfk_download_pressure.argtypes = [vp, vp, u64]
    L.lfk_upload_pressure.argtypes = [vp, vp, u64]
    L.lfk_apply_a.argtypes = [vp, db, vp, vp, u64]
    L.lfk_cfl.argtypes = [vp, C.POINTER(db)]
    L.lfk_time_step_cfl.argtypes = [vp, C.POINTER(db)]
    L.lfk_update.argtypes = [vp, db, C.POINTER(u64)]
    L.lfk_set_sources.argtypes = [vp, vp, u64]
    L.lfk_set_rng_seed.argtypes = [vp, u64]
    L.lfk_coerce_sources.argtypes = [vp]
    L.lfk_update_sources.argtypes = [vp, C.POINTER(u64)]
    L.lfk_mesher_sample.argtypes = [vp, C.POINTER(Mesher), db, vp, u64, vp]
    L.lfk_voxelize_mesh.argtypes = [vp, vp, u64, vp, u64, db, vp, vp, vp]
    L.lfk_voxels_download.argtypes = [vp, vp, u64]
    L.lfk_obstacle_cells.argtypes = [vp, vp, u64, C.POINTER(u64), ci]
    L.lfk_seed_box_device.argtypes = [vp, vp, vp, vp, C.c_uint32, u64, ci]
    L.lfk_synthetic_projection_device.argtypes = [vp, u64]
    L.lfk_set_timing.argtypes = [vp, ci]
    L.lfk_set_tuning.argtypes = [vp, C.c_char_p, ci]
    L.lfk_get_stats.argtypes = [vp, C.POINTER(Stats)]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data if isinstance(a, np.ndarray) else int(a))


def _v3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(3))


class PinnedBuffer:
    """page-locked host memory from lfk_host_alloc (numpy view in .array)"""

    def __init__(self, nbytes):
        L = load_library()
        self.L, self.ptr, self.nbytes = L, C.c_void_p(), int(nbytes)
        rc = L.lfk_host_alloc(C.byref(self.ptr), self.nbytes)
        if rc != 0:
            raise LfkError(rc, (L.lfk_last_error(None) or b"").decode())
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(max(self.nbytes, 1),))

    def close(self):
        if self.ptr:
            self.array = None
            self.L.lfk_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id():
    L = load_library()
    buf = (C.c_char * 128)()
    rc = L.lfk_nccl_unique_id(buf)
    if rc != 0:
        raise LfkError(rc, (L.lfk_last_error(None) or b"").decode())
    return bytes(buf)


class Context:
    """One lfk_ctx: one GPU, one z-slab of the grid."""

    def __init__(self, size, device=0, stream=None, nranks=1, rank=0, nccl_id=None, **params):
        self.L = load_library()
        self.size = tuple(int(s) for s in size)
        self.ptr = C.c_void_p()
        idbuf = C.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
        rc = self.L.lfk_create(C.byref(self.ptr), *self.size, int(device),
                               C.c_void_p(int(stream)) if stream else None, int(nranks), int(rank), idbuf)
        if rc != 0:
            self.ptr = None
            raise LfkError(rc, (self.L.lfk_last_error(None) or b"").decode())
        self.nranks, self.rank = int(nranks), int(rank)
        self.params = Params()
        self.L.lfk_get_params(self.ptr, C.byref(self.params))
        if params:
            self.set_params(**params)

    # -- plumbing --
    def close(self):
        if getattr(self, "ptr", None):
            self.L.lfk_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise LfkError(rc, (self.L.lfk_last_error(self.ptr) or b"").decode())

    def set_params(self, **kw):
        P = self.params
        alias = {"h": "cell_size", "offset": "grid_offset", "blend": "blending_factor", "skin": "boundary_skin_width",
                 "stiffness": "correction_stiffness", "extrap_iters": "extrapolation_iterations"}
        for k, v in kw.items():
            k = alias.get(k, k)
            if k in ("grid_offset", "gravity"):
                getattr(P, k)[:] = [float(x) for x in v]
            elif k in ("method", "extrapolation_iterations", "max_iterations", "preconditioner"):
                setattr(P, k, int(v))
            else:
                setattr(P, k, float(v))
        self._ck(self.L.lfk_set_params(self.ptr, C.byref(P)))

    @property
    def ncells(self):
        return self.size[0] * self.size[1] * self.size[2]

    def sync(self):
        self._ck(self.L.lfk_sync(self.ptr))

    def slab(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self.L.lfk_slab(self.ptr, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- state --
    def upload_particles(self, arr):
        if isinstance(arr, np.ndarray):
            arr = np.ascontiguousarray(arr, dtype=PARTICLE_DTYPE)
            self._keep = arr
            self._ck(self.L.lfk_upload_particles(self.ptr, _ptr(arr), arr.shape[0]))
        else:  # (raw host pointer, count), e.g. pinned memory owned by the caller
            ptr, n = arr
            self._ck(self.L.lfk_upload_particles(self.ptr, C.c_void_p(int(ptr)), int(n)))

    def num_particles(self):
        n = C.c_uint64()
        self._ck(self.L.lfk_num_particles(self.ptr, C.byref(n)))
        return n.value

    def download_particles(self, out=None):
        n = self.num_particles()
        if out is None:
            out = np.empty(n, dtype=PARTICLE_DTYPE)
        if isinstance(out, np.ndarray):
            got = C.c_uint64()
            self._ck(self.L.lfk_download_particles(self.ptr, _ptr(out), out.shape[0], C.byref(got)))
            return out[:got.value]
        ptr, cap = out
        got = C.c_uint64()
        self._ck(self.L.lfk_download_particles(self.ptr, C.c_void_p(int(ptr)), int(cap), C.byref(got)))
        return got.value

    def download_positions(self):
        n = self.num_particles()
        out = np.empty((n, 3), dtype=np.float64)
        got = C.c_uint64()
        self._ck(self.L.lfk_download_positions(self.ptr, _ptr(out), n, C.byref(got)))
        return out

    def download_positions_async(self, ptr, capacity):
        """queues the positions download into pinned memory at `ptr` (capacity in particles); returns the count"""
        got = C.c_uint64()
        self._ck(self.L.lfk_download_positions_async(self.ptr, C.c_void_p(int(ptr)), int(capacity), C.byref(got)))
        return got.value

    def wait_transfers(self):
        self._ck(self.L.lfk_wait_transfers(self.ptr))

    def checkpoint_save(self, path):
        self._ck(self.L.lfk_checkpoint_save(self.ptr, str(path).encode()))

    def checkpoint_load(self, path):
        self._ck(self.L.lfk_checkpoint_load(self.ptr, str(path).encode()))
        self.L.lfk_get_params(self.ptr, C.byref(self.params))

    def upload_cells(self, arr):
        arr = np.ascontiguousarray(arr, dtype=CELL_DTYPE)
        assert arr.shape[0] == self.ncells
        self._ck(self.L.lfk_upload_cells(self.ptr, _ptr(arr)))

    def download_cells(self, out=None):
        if out is None:
            out = np.zeros(self.ncells, dtype=CELL_DTYPE)
        self._ck(self.L.lfk_download_cells(self.ptr, _ptr(out)))
        return out

    def upload_cells_slab(self, ptr):
        """raw host pointer to the slab's layers [max(z0 - 1, 0), min(z1 + 1, nz)) (multi-GPU hosts, pinned memory)"""
        self._ck(self.L.lfk_upload_cells_slab(self.ptr, C.c_void_p(int(ptr))))

    def download_cells_slab(self, ptr):
        """raw host pointer to room for the owned layers [z0, z1)"""
        self._ck(self.L.lfk_download_cells_slab(self.ptr, C.c_void_p(int(ptr))))

    def upload_old_cells(self, arr):
        arr = np.ascontiguousarray(arr, dtype=CELL_DTYPE)
        self._ck(self.L.lfk_upload_old_cells(self.ptr, _ptr(arr)))

    def download_old_cells(self, out=None):
        if out is None:
            out = np.zeros(self.ncells, dtype=CELL_DTYPE)
        self._ck(self.L.lfk_download_old_cells(self.ptr, _ptr(out)))
        return out

    def download_table(self):
        b = np.zeros(self.ncells, dtype=np.uint64)
        c = np.zeros(self.ncells, dtype=np.uint64)
        self._ck(self.L.lfk_download_table(self.ptr, _ptr(b), _ptr(c)))
        return b, c

    def num_fluid_cells(self):
        n = C.c_uint64()
        self._ck(self.L.lfk_num_fluid_cells(self.ptr, C.byref(n)))
        return n.value

    def download_fluid_cells(self):
        n = self.num_fluid_cells()
        out = np.zeros(n, dtype=np.uint64)
        self._ck(self.L.lfk_download_fluid_cells(self.ptr, _ptr(out), n))
        return out

    # -- stages --
    def hash(self):
        self._ck(self.L.lfk_hash(self.ptr))

    def advect(self, dt):
        self._ck(self.L.lfk_advect(self.ptr, dt))

    def collide(self):
        self._ck(self.L.lfk_collide(self.ptr))

    def p2g(self):
        self._ck(self.L.lfk_p2g(self.ptr))

    def gravity(self, dt):
        self._ck(self.L.lfk_gravity(self.ptr, dt))

    def pressure_solve(self, dt):
        res, it = C.c_double(), C.c_uint64()
        self._ck(self.L.lfk_pressure_solve(self.ptr, dt, C.byref(res), C.byref(it)))
        return res.value, it.value

    def download_rhs(self, dt):
        n = self.num_fluid_cells()
        b = np.zeros(n, dtype=np.float64)
        f = np.zeros(n, dtype=np.uint8)
        self._ck(self.L.lfk_download_rhs(self.ptr, dt, _ptr(b), _ptr(f), n))
        return b, f

    def download_pressure(self):
        n = self.num_fluid_cells()
        p = np.zeros(n, dtype=np.float64)
        self._ck(self.L.lfk_download_pressure(self.ptr, _ptr(p), n))
        return p

    def upload_pressure(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        self._ck(self.L.lfk_upload_pressure(self.ptr, _ptr(p), p.shape[0]))

    def apply_a(self, dt, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.zeros_like(v)
        self._ck(self.L.lfk_apply_a(self.ptr, dt, _ptr(v), _ptr(out), v.shape[0]))
        return out

    def apply_pressure(self, dt):
        self._ck(self.L.lfk_apply_pressure(self.ptr, dt))

    def correct(self, dt):
        self._ck(self.L.lfk_correct(self.ptr, dt))

    def extrapolate(self):
        self._ck(self.L.lfk_extrapolate(self.ptr))

    def g2p(self):
        self._ck(self.L.lfk_g2p(self.ptr))

    def cfl(self):
        v = C.c_double()
        self._ck(self.L.lfk_cfl(self.ptr, C.byref(v)))
        return v.value

    def time_step(self, dt=None):
        if dt is None:
            used = C.c_double()
            self._ck(self.L.lfk_time_step_cfl(self.ptr, C.byref(used)))
            return used.value
        self._ck(self.L.lfk_time_step(self.ptr, dt))
        return dt

    def update(self, dt):
        n = C.c_uint64()
        self._ck(self.L.lfk_update(self.ptr, dt, C.byref(n)))
        return n.value

    def set_sources(self, sources, seed=None):
        """sources: list of dicts {cells: (n, 3) integer array of x, y, z, velocity, density (cubic root), active, coerce}"""
        arr = (Source * max(len(sources), 1))()
        keep = []
        for k, s in enumerate(sources):
            cells = np.ascontiguousarray(np.asarray(s["cells"], dtype=np.uint64).reshape(-1, 3))
            keep.append(cells)
            arr[k].cells = cells.ctypes.data
            arr[k].num_cells = cells.shape[0]
            arr[k].velocity[:] = [float(v) for v in s.get("velocity", (0, 0, 0))]
            arr[k].target_density_cubic_root = int(s.get("density", 2))
            arr[k].active = int(bool(s.get("active", True)))
            arr[k].coerce_velocity = int(bool(s.get("coerce", False)))
        self._ck(self.L.lfk_set_sources(self.ptr, C.byref(arr), len(sources)))
        if seed is not None:
            self._ck(self.L.lfk_set_rng_seed(self.ptr, int(seed)))

    def coerce_sources(self):
        self._ck(self.L.lfk_coerce_sources(self.ptr))

    def update_sources(self):
        n = C.c_uint64()
        self._ck(self.L.lfk_update_sources(self.ptr, C.byref(n)))
        return n.value

    def mesher_sample(self, size, offset, cell_size, extent, cell_radius, r, xyz=None):
        """mesher::_sample_surface_function on the device: (sz + 1, sy + 1, sx + 1) array; xyz None = own particles"""
        m = Mesher()
        m.grid_offset[:] = [float(v) for v in offset]
        m.cell_size, m.particle_extent, m.cell_radius = float(cell_size), float(extent), int(cell_radius)
        m.size[:] = [int(v) for v in size]
        out = np.zeros((int(size[2]) + 1, int(size[1]) + 1, int(size[0]) + 1), dtype=np.float64)
        if xyz is not None:
            xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        self._ck(self.L.lfk_mesher_sample(self.ptr, C.byref(m), float(r), _ptr(xyz), 0 if xyz is None else xyz.shape[0],
                                          _ptr(out)))
        return out

    def voxelize_mesh(self, positions, indices, cell_size, ref_offset):
        """voxelizer on the device: (grid_min (3,) int64, voxels (vz, vy, vx) u8)"""
        pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, dtype=np.uint64).ravel()
        off = _v3(ref_offset)
        gmin, vsz = np.zeros(3, dtype=np.int64), np.zeros(3, dtype=np.uint64)
        self._ck(self.L.lfk_voxelize_mesh(self.ptr, _ptr(pos), pos.shape[0], _ptr(idx), idx.shape[0], float(cell_size),
                                          _ptr(off), _ptr(gmin), _ptr(vsz)))
        vox = np.zeros((int(vsz[2]), int(vsz[1]), int(vsz[0])), dtype=np.uint8)
        self._ck(self.L.lfk_voxels_download(self.ptr, _ptr(vox), vox.size))
        return gmin, vox

    def obstacle_cells(self, mark_solid=False):
        n = C.c_uint64()
        self._ck(self.L.lfk_obstacle_cells(self.ptr, None, 0, C.byref(n), 0))
        cells = np.zeros((max(n.value, 1), 3), dtype=np.uint64)
        self._ck(self.L.lfk_obstacle_cells(self.ptr, _ptr(cells), n.value, C.byref(n), int(bool(mark_solid))))
        return cells[:n.value]

    def seed_box_device(self, start, size, velocity=(0, 0, 0), density=2, seed=1, append=False):
        a, b, v = _v3(start), _v3(size), _v3(velocity)  # keep the temporaries alive across the call
        self._ck(self.L.lfk_seed_box_device(self.ptr, _ptr(a), _ptr(b), _ptr(v), int(density), int(seed),
                                            int(append)))

    def synthetic_projection_device(self, seed=1):
        self._ck(self.L.lfk_synthetic_projection_device(self.ptr, int(seed)))

    # -- instrumentation --
    def set_timing(self, on):
        self._ck(self.L.lfk_set_timing(self.ptr, int(bool(on))))

    def set_tuning(self, key, value):
        """A/B switch between kernel variants computing the same result (lfk_set_tuning)."""
        self._ck(self.L.lfk_set_tuning(self.ptr, key.encode(), int(value)))

    def stats(self):
        s = Stats()
        self._ck(self.L.lfk_get_stats(self.ptr, C.byref(s)))
        d = dict(kernel_launches=s.kernel_launches, pcg_iterations=s.pcg_iterations, pcg_residual=s.pcg_residual,
                 num_particles=s.num_particles, num_fluid_cells=s.num_fluid_cells,
                 exchanged_particles=s.exchanged_particles)
        d["phase_ms"] = {PHASES[i]: s.phase_ms[i] for i in range(len(PHASES))}
        return d

    def reset_stats(self):
        self._ck(self.L.lfk_reset_stats(self.ptr))
