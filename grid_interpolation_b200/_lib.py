"""ctypes binding of libgridinterp.so (include/gridinterp.h).  Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GI_LIBRARY: another build of the same library (A/B variants from `python -m grid_interpolation_b200.build --variant`)
LIB_PATH = os.environ.get("GI_LIBRARY") or os.path.join(_HERE, "libgridinterp.so")

GI_OK, GI_ERR_BAD_ARG, GI_ERR_NO_DEVICE, GI_ERR_OOM, GI_ERR_CUDA = 0, 1, 2, 3, 4
GI_ERR_ITER_CAP, GI_ERR_TOO_FEW_POINTS, GI_ERR_BAD_INDEX, GI_ERR_HALO = 5, 6, 7, 8
GI_ERR_PACK_WINDOW, GI_ERR_TIMEOUT = 9, 10
GI_HOST, GI_DEVICE = 0, 1
GI_POINTS_ROWMAJOR, GI_TOPO_ROWMAJOR, GI_FIELD_ROWMAJOR, GI_KD_REFERENCE_VISITS = 1, 2, 4, 8
GI_ESRG_REFERENCE_ORDER = 16
GI_KD_EXACT_PRUNE = 32
GI_BILINEAR_SEARCH_AXES = 64
GI_KD_FAITHFUL_BUILD = 128
GI_BARY_REFERENCE_ORDER = 256
GI_BARY_BAND_PACK = 512
GI_BARY_PLANE_VALUES = 2048
GI_TARGETS_ROWMAJOR = 1024
GI_MAX_GATHER_FIELDS, GI_PEER_HANDLE_BYTES = 8, 64

_vp, _i, _i64, _d, _f = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_float

# name -> argtypes; every symbol include/gridinterp.h declares (tests/test_abi.py checks the two agree)
SIGNATURES = {
    "gi_version": ([], _i),
    "gi_last_error": ([], C.c_char_p),
    "gi_device_count": ([], _i),
    "gi_set_device": ([_i], _i),
    "gi_host_alloc": ([C.c_uint64], _vp),
    "gi_host_free": ([_vp], None),
    "gi_stream_status": ([_vp], _i),
    "gi_set_profiling": ([_i], None),
    "gi_set_pipeline_band_bytes": ([C.c_uint64], None),
    "gi_last_timings": ([_vp, _i], _i),
    "gi_last_timing_name": ([_i], C.c_char_p),
    "gi_release_workspace": ([], None),
    "gi_set_fixup_list_capacity": ([_i64], None),
    "gi_set_sliver_guard": ([C.c_double], None),
    "gi_peer_alloc": ([C.c_uint64], _vp),
    "gi_peer_free": ([_vp], None),
    "gi_peer_export": ([_vp, _vp], _i),
    "gi_peer_open": ([_vp, _vp], _i),
    "gi_peer_close": ([_vp], _i),
    "gi_peer_signal": ([_vp, _i, _i, C.c_uint32, _vp], _i),
    "gi_peer_wait": ([_vp, _i, C.c_uint32, _d, _vp], _i),
    "gi_plan_destroy": ([_vp], None),
}
for _suf, _real in (("f64", _d), ("f32", _f)):
    SIGNATURES.update({
        f"gi_bilinear_{_suf}": ([_vp, _i, _vp, _i, _vp, _vp, _i64, _vp, _i, _i, _vp], _i),
        f"gi_idw_kdtree_{_suf}": ([_vp, _i, _vp, _vp, _i, _vp, _i, _i, _vp, _i, _i, _vp], _i),
        f"gi_idw_esrg_nearest_{_suf}": ([_vp, _i, _vp, _vp, _i, _vp, _i, _real, _i, _vp, _i, _i, _vp], _i),
        f"gi_barycentric_interpolation_{_suf}": ([_vp, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _vp], _i),
        f"gi_barycentric_interpolation_ex_{_suf}": ([_vp, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i,
                                                     _vp, _vp, _vp, _vp, _vp, _i, _i, _vp], _i),
        f"gi_barycentric_interpolation_gather_{_suf}": ([_vp, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i,
                                                         _vp, _i, _i, _vp], _i),
        f"gi_idw_kdtree_ex_{_suf}": ([_vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp], _i),
        f"gi_idw_esrg_nearest_ex_{_suf}": ([_vp, _i, _vp, _vp, _i, _vp, _i, _real, _i, _i, _i, _vp, _vp, _vp, _vp,
                                            _i, _i, _vp], _i),
        f"gi_barycentric_plan_create_{_suf}": ([_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp], _i),
        f"gi_idw_kdtree_plan_create_{_suf}": ([_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp], _i),
        f"gi_idw_esrg_plan_create_{_suf}": ([_vp, _i, _vp, _i, _vp, _i, _real, _i, _i, _i, _i, _i, _vp, _vp], _i),
        f"gi_plan_execute_{_suf}": ([_vp, _vp, _vp, _i, _vp], _i),
        f"gi_kd_build_{_suf}": ([_vp, _i, _vp, _i, _i, _vp], _i),
        f"gi_esrg_assign_points_to_cells_{_suf}": ([_vp, _i, _vp, _i, _vp, _i, _real, _vp, _vp, _i, _i, _vp], _i),
        f"gi_fill_nearest_neighbour_{_suf}": ([_vp, _i, _i, _vp, _vp, _i, _i, _vp], _i),
        f"gi_bilinear_morton_order_{_suf}": ([_vp, _i, _vp, _i, _vp, _i64, _vp, _i, _i, _vp], _i),
        f"gi_barycentric_points_{_suf}": ([_vp, _i, _vp, _vp, _vp, _i, _vp, _i64, _vp, _vp, _vp, _i, _i, _vp], _i),
        f"gi_idw_esrg_nearest_points_{_suf}": ([_vp, _i, _vp, _vp, _i, _vp, _i, _real, _vp, _i64, _i, _vp, _vp, _vp, _i, _i, _vp], _i),
        f"gi_mesh_reorder_{_suf}": ([_vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp], _i),
        f"gi_idw_kdtree_points_{_suf}": ([_vp, _i, _vp, _vp, _i64, _i, _vp, _vp, _vp, _i, _i, _vp], _i),
    })

_lib = None


def lib():
    """Load libgridinterp.so (building it with nvcc if the in-tree binary is missing or stale)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or os.environ.get("GI_REBUILD"):
            from . import build as _build
            _build.build()
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing and could not be built: the CUDA extension is required "
                              "(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError = ABI mismatch: fail loudly
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().gi_last_error().decode()


class GiError(RuntimeError):
    """A device-side / runtime failure reported by libgridinterp; ``status`` is the gi_status code."""

    def __init__(self, msg, status):
        super().__init__(msg)
        self.status = status


def check(status: int, what: str) -> None:
    if status == GI_OK:
        return
    msg = f"{what}: {last_error()} (gi_status {status})"
    if status == GI_ERR_BAD_ARG:
        raise ValueError(msg)
    if status == GI_ERR_OOM:
        raise MemoryError(msg)
    raise GiError(msg, status)
