"""CPU-side checks of the C-ABI boundary: the library builds, loads, exports every symbol include/lfk.h declares,
and refuses to run without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from libfluid_b200 import build as lfk_build
from libfluid_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    lfk_build.build()
    return capi.load_library()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "lfk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lfk_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.lfk_abi_version() == 1


def test_struct_layouts_match_header():
    # lfk_params: 3+1+1+3+1+1+1+1+1 doubles, then 4 int32 ; lfk_stats: u64,u64,f64,16 f64,u64,u64,u64
    assert C.sizeof(capi.Params) == 13 * 8 + 4 * 4
    assert C.sizeof(capi.Stats) == (3 + 16 + 3) * 8
    assert capi.PARTICLE_DTYPE.itemsize == 152 and capi.CELL_DTYPE.itemsize == 32


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.LfkError) as ei:
        capi.Context((8, 8, 8), cell_size=1.0)
    assert ei.value.code == -2001  # LFK_E_NO_DEVICE
    assert b"no CPU fallback" in lib.lfk_last_error(None)


def test_create_rejects_bad_arguments(lib):
    ptr = C.c_void_p()
    assert lib.lfk_create(C.byref(ptr), 0, 8, 8, 0, None, 1, 0, None) == -2000
    assert lib.lfk_create(C.byref(ptr), 8, 8, 8, 0, None, 4, 0, None) == -2000   # slabs thinner than 4 cells
    assert lib.lfk_create(None, 8, 8, 8, 0, None, 1, 0, None) == -2000
    assert lib.lfk_set_params(None, None) == -2000
