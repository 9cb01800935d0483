"""The C-ABI library loads and exports exactly what include/cs_b200.h declares
(no compute calls -- runs without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import circuitscape_b200 as cb
from circuitscape_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "cs_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cs_b200_[a-z0-9_]+)\s*\(", txt)))


def test_library_built():
    assert os.path.isfile(_lib.LIB_PATH), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/cs_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms, "python binding and header disagree"


def test_version():
    assert _lib.load().cs_b200_version() >= 1000


def test_opts_struct_layout():
    # sizeof must match the C struct: 4 int32 + 2 double + 8 int32
    assert ctypes.sizeof(_lib.Opts) == 4 * 4 + 2 * 8 + 8 * 4
    assert ctypes.sizeof(_lib.Stats) == 8 * 8


def test_bad_arguments_are_rejected_without_a_device():
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.cs_b200_create(0, 0, None, None, None, 32, 0, 1, 0, None, ctypes.byref(h))
    assert rc == _lib.ERR_ARG
    assert b"bad matrix" in lib.cs_b200_last_error(None)


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cb.B200Unavailable):
        cb.B200Factor(sp.identity(4, format="csr"), cb.CUDASolver())
    with pytest.raises(cb.B200Unavailable):
        cb.B200Factor.from_raster(np.ones((3, 3)), cb.CUDASolver())


def test_from_raster_rejects_bad_arguments_without_a_device():
    lib = _lib.load()
    h = ctypes.c_void_p()
    g = np.ones((2, 2))
    rc = lib.cs_b200_create_from_raster(0, 2, g.ctypes.data, 1, 0, 0, 0, None, ctypes.byref(h), None, None)
    assert rc == _lib.ERR_ARG
    rc = lib.cs_b200_create_from_raster(20000, 20000, g.ctypes.data, 1, 0, 0, 0, None, ctypes.byref(h), None, None)
    assert rc != 0 and b"too large" in lib.cs_b200_last_error(None)
