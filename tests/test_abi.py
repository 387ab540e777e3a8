"""CPU tests of the C-ABI boundary: the library loads, exports every declared symbol, validates its
arguments and agrees with the oracle on the level layout.  No compute kernels are launched."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nvp_b200 import _lib
from oracle import nvp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "nvp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nvp_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = declared_symbols()
    assert set(syms) == set(_lib.EXPORTS), (syms, _lib.EXPORTS)
    for s in syms:
        assert hasattr(lib, s), f"libnvp_b200.so does not export {s}"
    assert lib.nvp_version() >= 100


def test_struct_layout_matches_header():
    assert C.sizeof(_lib.NvpDesc) == 11 * 4
    assert C.sizeof(_lib.NvpPtrs) == 18 * 8


@pytest.mark.parametrize("levels,base,pls", [(16, 16, 1.35), (8, 4, 2.0), (5, 16, 1.5), (1, 7, 1.35)])
def test_level_table_bit_equal_to_oracle(levels, base, pls):
    d = _lib.NvpDesc(2, levels, base, pls, 2, 6, 20, 24, 128, 3, 30.0)
    sc, rs, of = _lib.level_table(d)
    t = O.level_table(levels, base, pls)
    assert np.array_equal(np.asarray(sc, np.float32).view(np.uint32), t.scales.view(np.uint32))
    assert rs == list(t.res) and of == list(t.offsets)


def test_latent_dim_and_desc_validation():
    lib = _lib.load()
    d = _lib.NvpDesc(2, 16, 16, 1.35, 2, 600, 300, 300, 128, 3, 30.0)
    assert lib.nvp_latent_dim(C.byref(d)) == 114
    d4 = _lib.NvpDesc(4, 16, 16, 1.35, 4, 600, 300, 300, 128, 3, 30.0)
    assert lib.nvp_latent_dim(C.byref(d4)) == 228
    out = C.c_size_t(0)
    assert lib.nvp_workspace_bytes(C.byref(d), 1 << 16, _lib.MODE_FP32_SIMT, 1, C.byref(out)) == 0 and out.value > 0
    bad = _lib.NvpDesc(3, 16, 16, 1.35, 2, 600, 300, 300, 128, 3, 30.0)
    assert lib.nvp_workspace_bytes(C.byref(bad), 16, 0, 0, C.byref(out)) != 0
    assert b"n_features_per_level" in lib.nvp_last_error()
    bad = _lib.NvpDesc(2, 16, 16, 1.35, 2, 600, 300, 300, 64, 3, 30.0)
    assert lib.nvp_workspace_bytes(C.byref(bad), 16, 0, 0, C.byref(out)) != 0
    bad = _lib.NvpDesc(2, 40, 16, 1.35, 2, 600, 300, 300, 128, 3, 30.0)
    with pytest.raises(RuntimeError):
        _lib.level_table(bad)


def test_compute_entry_points_reject_null_and_host_pointers():
    lib = _lib.load()
    d = _lib.NvpDesc(2, 16, 16, 1.35, 2, 6, 20, 24, 128, 3, 30.0)
    pp = _lib.NvpPtrs()
    host = (C.c_float * 16)()
    rc = lib.nvp_forward(C.byref(d), C.byref(pp), C.addressof(host), C.addressof(host), 4, C.addressof(host),
                         C.addressof(host), 64, 0, None)
    assert rc != 0 and len(lib.nvp_last_error()) > 0
    # n == 0 is a no-op, not an error (empty batch)
    assert lib.nvp_forward(C.byref(d), C.byref(pp), None, None, 0, None, None, 0, 0, None) == 0
