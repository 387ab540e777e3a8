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


@pytest.mark.parametrize("F,n", [(2, 1245184), (4, 1245184), (2, 5000), (1, 100000), (8, 65536)])
def test_grid_bin_plan_windows_cover_every_in_range_sample(F, n):
    """Host logic of the tile-binned grid path (csrc/grid_binned.cuh), checked against the oracle's index arithmetic with no
    GPU: for every coordinate in [0,1] - random, k/TB tile boundaries and their neighbours in fp32, the pixel / frame
    lattices of the BASELINE videos - the sample's cell and its +1 corner lie inside the window of its tile at every level
    (cell(u) - cell(tile origin) in [0, min(E-2, res-1-lo)]), so the kernels' direct-access fallback is only ever taken by
    out-of-range coordinates."""
    d = _lib.NvpDesc(F, 16, 16, 1.35, F, 600, 300, 300, 128, 3, 30.0)
    plan = _lib.grid_bin_plan(d, n)
    assert plan, "both reference configurations must use the binned path"
    tb, ext, base = plan["tiles_per_axis"], plan["window_extent"], plan["window_base"]
    assert tb & (tb - 1) == 0 and plan["chunk"] % 32 == 0 and plan["chunk"] >= 32
    assert base[0] == 0 and all(base[l + 1] - base[l] == ext[l] ** 2 for l in range(16))
    assert plan["workspace"] >= 3 * n * 16 + 3 * tb * tb * 12
    assert base[16] * F * 4 <= 216 * 1024, "one window region must fit an SM's shared memory"
    t = O.level_table(16, 16, 1.35)
    rng = np.random.default_rng(F * 1000 + n % 997)
    k = np.arange(tb + 1, dtype=np.float32) / np.float32(tb)
    u = np.concatenate([rng.random(400000).astype(np.float32), k, np.nextafter(k, np.float32(0)), np.nextafter(k, np.float32(2)),
                        np.arange(600, dtype=np.float32) / np.float32(599), np.arange(1080, dtype=np.float32) / np.float32(1079),
                        np.arange(1920, dtype=np.float32) / np.float32(1919), np.arange(300, dtype=np.float32) / np.float32(299)])
    u = u[(u >= 0) & (u <= 1)]
    b = np.clip((u * np.float32(tb)).astype(np.int32), 0, tb - 1)       # bin_tile_axis: exact product, truncation
    ub = b.astype(np.float32) / np.float32(tb)
    assert np.all(ub <= u)
    for l in range(16):
        s, res, E = t.scales[l], int(t.res[l]), ext[l]
        assert 3 <= E <= res + 1
        cell = np.floor(O._fmaf(np.full_like(u, s), u, 0.5)).astype(np.int64)
        lo = np.floor(O._fmaf(np.full_like(u, s), ub, 0.5)).astype(np.int64)
        aa = cell - lo
        amax = np.minimum(E - 2, res - 1 - lo)
        assert aa.min() >= 0 and np.all(aa <= amax), (l, E, int(aa.max()), int((aa > amax).sum()))
        # the reciprocal trick of region_io: idx // E == (idx * magic) >> 20 on the whole window
        idx = np.arange(E * E, dtype=np.uint64)
        magic = np.uint64((1 << 20) // E + 1)
        assert np.array_equal((idx * magic) >> np.uint64(20), idx // np.uint64(E))


def test_grid_bin_plan_is_off_for_configurations_the_binned_path_does_not_serve():
    # 5 levels x 1 feature = 5 latent columns per plane: not a whole number of 16-byte chunks -> direct kernels
    assert _lib.grid_bin_plan(_lib.NvpDesc(1, 5, 16, 1.5, 1, 6, 20, 24, 128, 3, 30.0), 4096) == {}
    # more samples than 32-bit row offsets into the latent tile buffer can address
    assert _lib.grid_bin_plan(_lib.NvpDesc(2, 16, 16, 1.35, 2, 600, 300, 300, 128, 3, 30.0), 9_000_000) == {}
    os.environ["NVP_GRID_BINNED"] = "0"
    try:
        assert _lib.grid_bin_plan(_lib.NvpDesc(2, 16, 16, 1.35, 2, 600, 300, 300, 128, 3, 30.0), 4096) == {}
    finally:
        del os.environ["NVP_GRID_BINNED"]
