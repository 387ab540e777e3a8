"""ctypes binding of libnvp_b200.so (the C ABI declared in include/nvp_b200.h).

The product path has no CPU or eager fallback: if the shared library is missing this module raises
at first use, and every compute entry point rejects host pointers.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# NVP_B200_LIB: load another build of the library (development: ablation builds of scripts/fused_ablate.sh)
LIB_PATH = os.environ.get("NVP_B200_LIB") or os.path.join(_HERE, "libnvp_b200.so")

NVP_MAX_LEVELS = 32
NVP_MAX_LAYERS = 3
MODE_FP32_SIMT = 0
MODE_TC_F16 = 1
FLAG_TEMPORAL_INTERP = 0x100
MODES = {"fp32": MODE_FP32_SIMT, "simt": MODE_FP32_SIMT, "tc": MODE_TC_F16, "f16": MODE_TC_F16}


class NvpDesc(C.Structure):
    _fields_ = [
        ("n_features", C.c_int32), ("n_levels", C.c_int32), ("base_resolution", C.c_int32),
        ("per_level_scale", C.c_float), ("sparse_features", C.c_int32), ("t_resolution", C.c_int32),
        ("x_resolution", C.c_int32), ("y_resolution", C.c_int32), ("hidden", C.c_int32),
        ("n_layers", C.c_int32), ("w0_first", C.c_float),
    ]


class NvpPtrs(C.Structure):
    """nvp_params and nvp_grads share this field order."""
    _fields_ = [
        ("kf_xy", C.c_void_p), ("kf_yt", C.c_void_p), ("kf_xt", C.c_void_p), ("sparse", C.c_void_p),
        ("siren_w", C.c_void_p * NVP_MAX_LAYERS), ("siren_b", C.c_void_p * NVP_MAX_LAYERS),
        ("last_w", C.c_void_p), ("last_b", C.c_void_p),
        ("mod_w", C.c_void_p * NVP_MAX_LAYERS), ("mod_b", C.c_void_p * NVP_MAX_LAYERS),
    ]


EXPORTS = (
    "nvp_version", "nvp_last_error", "nvp_level_table", "nvp_latent_dim", "nvp_workspace_bytes",
    "nvp_encode_latent", "nvp_forward", "nvp_backward", "nvp_fwd_loss_bwd", "nvp_last_launch_count",
    "nvp_selftest_umma", "nvp_profile_enable", "nvp_profile_read", "nvp_adamw_step", "nvp_sample_batch", "nvp_scatter_latent",
    "nvp_record_grid_grads_event", "nvp_grid_bin_plan", "nvp_debug_timeline_read",
)

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run `make` (or "
            "`python -c 'import __graft_entry__ as g; g.build()'`) in the repo root. There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    D, P = C.POINTER(NvpDesc), C.POINTER(NvpPtrs)
    lib.nvp_version.restype = i32
    lib.nvp_last_error.restype = C.c_char_p
    lib.nvp_last_launch_count.restype = i32
    lib.nvp_latent_dim.argtypes = [D]
    lib.nvp_level_table.argtypes = [D, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    lib.nvp_workspace_bytes.argtypes = [D, i64, i32, i32, C.POINTER(C.c_size_t)]
    lib.nvp_encode_latent.argtypes = [D, P, vp, i64, vp, vp]
    lib.nvp_scatter_latent.argtypes = [D, vp, i64, vp, P, vp]
    lib.nvp_forward.argtypes = [D, P, vp, vp, i64, vp, vp, C.c_size_t, i32, vp]
    lib.nvp_backward.argtypes = [D, P, vp, vp, vp, i64, P, vp, C.c_size_t, i32, vp]
    lib.nvp_fwd_loss_bwd.argtypes = [D, P, vp, vp, vp, i64, i64, P, vp, vp, vp, C.c_size_t, i32, vp]
    lib.nvp_adamw_step.argtypes = [vp, vp, vp, vp, i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, i64, i32, vp]
    lib.nvp_sample_batch.argtypes = [vp, i32, i32, i32, vp, vp, i64, vp, vp, C.c_uint64, C.c_uint64, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.nvp_record_grid_grads_event.argtypes = [vp]
    lib.nvp_debug_timeline_read.argtypes = [C.POINTER(C.c_uint64), C.c_int32]
    lib.nvp_grid_bin_plan.argtypes = [D, i64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_size_t)]
    lib.nvp_profile_enable.argtypes = [i32]
    lib.nvp_profile_read.argtypes = [i32, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    lib.nvp_selftest_umma.argtypes = [vp, vp, vp, i32, i32, vp]
    for name in EXPORTS:
        if name not in ("nvp_last_error",):
            getattr(lib, name).restype = i32
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().nvp_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def make_desc(cfg: dict) -> NvpDesc:
    """cfg = config_nvp_*.json["nvp"].  The three 2-D encodings must agree (they do in both configs)."""
    exy, ext, eyt = cfg["2d_encoding_xy"], cfg["2d_encoding_xt"], cfg["2d_encoding_yt"]
    for k in ("n_levels", "n_features_per_level", "base_resolution", "per_level_scale"):
        if not (exy[k] == ext[k] == eyt[k]):
            raise ValueError(f"2d_encoding_xy/xt/yt must share {k}")
    if exy.get("otype", "DenseGrid") != "DenseGrid":
        raise ValueError("only otype=DenseGrid is supported for the 2-D encodings")
    s, n = cfg["3d_encoding"], cfg["network"]
    if s.get("upsample", False):
        raise NotImplementedError("3d_encoding.upsample=true is disabled in both reference configs and not built")
    return NvpDesc(exy["n_features_per_level"], exy["n_levels"], exy["base_resolution"], exy["per_level_scale"],
                   s["n_features_per_level"], s["t_resolution"], s["x_resolution"], s["y_resolution"],
                   n["n_neurons"], n["n_hidden_layers"], 30.0)


def level_table(desc: NvpDesc):
    """-> (scales [L] float, res [L] int, offsets [L+1] int) from the library."""
    L = desc.n_levels
    sc = (C.c_float * L)()
    rs = (C.c_int32 * L)()
    of = (C.c_int64 * (L + 1))()
    check(load().nvp_level_table(C.byref(desc), sc, rs, of), "nvp_level_table")
    return list(sc), list(rs), list(of)


def workspace_bytes(desc: NvpDesc, n: int, mode: int, what: int) -> int:
    out = C.c_size_t(0)
    check(load().nvp_workspace_bytes(C.byref(desc), n, mode, what, C.byref(out)), "nvp_workspace_bytes")
    return int(out.value)


def grid_bin_plan(desc: NvpDesc, n: int) -> dict:
    """Host-only: the tile-binned grid plan used for n samples ({} when the direct kernels are used instead)."""
    L = desc.n_levels
    tb, chunk, ws = C.c_int32(0), C.c_int32(0), C.c_size_t(0)
    ext, base = (C.c_int32 * L)(), (C.c_int32 * (L + 1))()
    check(load().nvp_grid_bin_plan(C.byref(desc), n, C.byref(tb), C.byref(chunk), ext, base, C.byref(ws)), "nvp_grid_bin_plan")
    if tb.value == 0:
        return {}
    return {"tiles_per_axis": tb.value, "chunk": chunk.value, "window_extent": list(ext), "window_base": list(base),
            "workspace": int(ws.value)}


PROFILE_KINDS = ("pack", "grid_gather", "mlp_forward", "mlp_backward", "mlp_wgrad", "grid_scatter", "fp32_mode", "misc", "grid_bin",
                 "mlp_fused")


def profile_enable(on: bool) -> None:
    check(load().nvp_profile_enable(1 if on else 0), "nvp_profile_enable")


def profile_read():
    """-> {kind: (total_ms, launches)} since the last read; synchronises the recorded events."""
    n = len(PROFILE_KINDS)
    ms = (C.c_float * n)()
    cnt = (C.c_int * n)()
    check(load().nvp_profile_read(n, ms, cnt), "nvp_profile_read")
    return {PROFILE_KINDS[i]: (float(ms[i]), int(cnt[i])) for i in range(n)}
