"""ctypes handle on libfluid_host.so's extern "C" test hooks (libfluid_b200/host/test_hooks.cpp): the C++ mirror of
fluid::simulation driven from Python (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

from libfluid_b200 import build as lfk_build
from libfluid_b200 import capi

HOST_DIR = os.path.join(os.path.dirname(os.path.abspath(capi.__file__)), "host")
HOST_SO = os.path.join(os.path.dirname(capi.LIB_PATH), "libfluid_host.so")


class CbLog(C.Structure):
    _fields_ = [("calls", C.c_int * 8), ("residual", C.c_double), ("max_pressure", C.c_double),
                ("iterations", C.c_size_t), ("pressure_len", C.c_size_t)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        lfk_build.build()
        subprocess.check_call(["make", "-s", "-C", HOST_DIR])
        L = C.CDLL(HOST_SO)
        vp, sz, db = C.c_void_p, C.c_size_t, C.c_double
        L.hapi_create.restype = vp
        L.hapi_create.argtypes = [sz, sz, sz, db, vp, vp, C.c_int, db]
        L.hapi_destroy.argtypes = [vp]
        L.hapi_seed_box.argtypes = [vp, vp, vp, sz]
        L.hapi_seed_sphere.argtypes = [vp, vp, db, sz]
        L.hapi_set_solid.argtypes = [vp, vp]
        L.hapi_add_source.argtypes = [vp, vp, sz, vp, C.c_int]
        L.hapi_num_particles.restype = sz
        L.hapi_num_particles.argtypes = [vp]
        for n in ("hapi_get_particles", "hapi_get_cells", "hapi_set_cells"):
            getattr(L, n).argtypes = [vp, vp]
        L.hapi_set_particles.argtypes = [vp, vp, sz]
        L.hapi_time_step.argtypes = [vp, db, vp, sz]
        L.hapi_update.argtypes = [vp, db, vp, sz]
        L.hapi_install_callbacks.argtypes = [vp, C.POINTER(CbLog)]
        L.hapi_cfl.restype = db
        L.hapi_cfl.argtypes = [vp]
        L.hapi_last_solve.argtypes = [vp, C.POINTER(db), C.POINTER(sz)]
        _lib = L
    return _lib


def _v3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(3))


class HostSim:
    def __init__(self, size, h=1.0, offset=(0, 0, 0), gravity=(0, -981.0, 0), method=capi.APIC, blend=1.0):
        self.L = lib()
        self.size = tuple(int(s) for s in size)
        o, g = _v3(offset), _v3(gravity)
        self.ptr = self.L.hapi_create(*self.size, float(h), o.ctypes.data, g.ctypes.data, int(method), float(blend))
        self.log = None

    def __del__(self):
        if getattr(self, "ptr", None):
            self.L.hapi_destroy(self.ptr)
            self.ptr = None

    def seed_box(self, start, size, dens=2):
        a, b = _v3(start), _v3(size)
        self.L.hapi_seed_box(self.ptr, a.ctypes.data, b.ctypes.data, dens)

    def seed_sphere(self, center, radius, dens=2):
        c = _v3(center)
        self.L.hapi_seed_sphere(self.ptr, c.ctypes.data, float(radius), dens)

    def set_solid(self, mask_zyx):
        m = np.ascontiguousarray(np.asarray(mask_zyx).reshape(-1), dtype=np.uint8)
        self.L.hapi_set_solid(self.ptr, m.ctypes.data)

    def add_source(self, cells_xyz, vel, coerce=False):
        a = np.ascontiguousarray(np.asarray(cells_xyz, dtype=np.uint64).reshape(-1, 3))
        v = _v3(vel)
        self.L.hapi_add_source(self.ptr, a.ctypes.data, a.shape[0], v.ctypes.data, int(coerce))

    def particles(self):
        out = np.empty(self.L.hapi_num_particles(self.ptr), dtype=capi.PARTICLE_DTYPE)
        self.L.hapi_get_particles(self.ptr, out.ctypes.data)
        return out

    def set_particles(self, arr):
        arr = np.ascontiguousarray(arr, dtype=capi.PARTICLE_DTYPE)
        self.L.hapi_set_particles(self.ptr, arr.ctypes.data, arr.shape[0])

    def cells(self):
        out = np.zeros(self.size[0] * self.size[1] * self.size[2], dtype=capi.CELL_DTYPE)
        self.L.hapi_get_cells(self.ptr, out.ctypes.data)
        return out

    def set_cells(self, arr):
        arr = np.ascontiguousarray(arr, dtype=capi.CELL_DTYPE)
        self.L.hapi_set_cells(self.ptr, arr.ctypes.data)

    def time_step(self, dt=None):
        err = C.create_string_buffer(512)
        rc = self.L.hapi_time_step(self.ptr, -1.0 if dt is None else float(dt), err, 512)
        if rc != 0:
            raise RuntimeError(err.value.decode())

    def update(self, dt):
        err = C.create_string_buffer(512)
        if self.L.hapi_update(self.ptr, float(dt), err, 512) != 0:
            raise RuntimeError(err.value.decode())

    def install_callbacks(self):
        self.log = CbLog()
        self.L.hapi_install_callbacks(self.ptr, C.byref(self.log))
        return self.log

    def cfl(self):
        return self.L.hapi_cfl(self.ptr)

    def last_solve(self):
        r, i = C.c_double(), C.c_size_t()
        self.L.hapi_last_solve(self.ptr, C.byref(r), C.byref(i))
        return r.value, i.value
