"""
ctypes binding of libglassb200.so (the C ABI in include/glass_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails the
functions here raise.  PyTorch is used only for device memory and streams.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libglassb200.so"
_lib = None

# status codes (include/glass_b200.h)
GLB_OK = 0
GLB_ERR_INVALID_ARG = -1
GLB_ERR_UNSUPPORTED = -2
GLB_ERR_CUDA = -3
GLB_ERR_NOMEM = -4
GLB_ERR_NOT_POSDEF = -10
GLB_ERR_NEGATIVE_CL = -11

T_NORMAL, T_LOGNORMAL, T_SQUARED_NORMAL = 0, 1, 2

# every symbol include/glass_b200.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _dp, _ip = C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p
SIGNATURES = {
    "glb_version": (C.c_char_p, []),
    "glb_last_error": (C.c_char_p, []),
    "glb_status_string": (C.c_char_p, [_i]),
    "glb_plan_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i]),
    "glb_plan_destroy": (_i, [_vp]),
    "glb_plan_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i), C.POINTER(_i64)]),
    "glb_alm2map": (_i, [_vp, _dp, _i, _dp, _ip, _dp, _vp]),
    "glb_alm2map_spin": (_i, [_vp, _dp, _dp, _i, _dp, _dp, _vp]),
    "glb_alm2map_spin_batch": (_i, [_vp, _dp, _i, _i, _dp, _dp, _vp]),
    "glb_map2alm": (_i, [_vp, _dp, _dp, _i, _dp, _vp]),
    "glb_map2alm_batch": (_i, [_vp, _dp, _i, _dp, _i, _dp, _vp]),
    "glb_almxfl": (_i, [_i, _dp, _dp, _i, _vp]),
    "glb_alm_draw": (_i, [_i, C.c_uint64, C.c_uint32, _dp, _vp]),
    "glb_alm_glass_to_healpix": (_i, [_i, _dp, _dp, _vp]),
    "glb_alm_combine": (_i, [_i, _i, _vp, _dp, _i, _dp, _vp]),
    "glb_iternorm_step": (_i, [_i, _i, _i, _dp, _dp, _dp, _dp, _dp, _dp, _vp, _vp]),
    "glb_points_workspace_bytes": (C.c_size_t, [_i64]),
    "glb_points_counts": (_i, [_i64, _dp, _dp, _i, C.c_double, C.c_double, _i, _dp, C.c_uint64, C.c_uint32, _dp, _dp, _dp, _dp, _i64, _dp, _vp, _vp]),
    "glb_points_cuts": (_i, [_dp, _i64, _i64, _i64, _i64, _i, _dp, _dp, _vp]),
    "glb_points_cuts_list": (_i, [_dp, _i64, _i64, _i64, _i64, _i64, _i, _dp, _dp, _vp]),
    "glb_points_fill_list": (_i, [_i64, _dp, _i64, _i64, _dp, _dp, C.c_uint64, C.c_uint32, _dp, _dp, _vp]),
    "glb_points_fill": (_i, [_i64, _dp, _dp, _i64, _i64, _dp, _dp, C.c_uint64, C.c_uint32, _dp, _dp, _dp, _vp]),
    "glb_ring2ang_uv": (_i, [_i64, _dp, _dp, _dp, _i64, _i, _dp, _dp, _vp]),
    "glb_randang": (_i, [_i64, _dp, _i64, C.c_uint64, C.c_uint32, _i, _dp, _dp, _vp]),
    "glb_uniform_positions": (_i, [_i64, _dp, _dp, C.c_uint64, C.c_uint32, _dp, _dp, _vp]),
    "glb_ang2pix": (_i, [_i64, _dp, _dp, _i64, _i, _dp, _vp]),
    "glb_multiplane_update": (_i, [_dp, _dp, _dp, C.c_double, _i64, C.c_double, C.c_double, _vp]),
    "glb_displace": (_i, [_dp, _dp, _dp, _dp, _i64, _i, _i64, _dp, _dp, _vp]),
    "glb_displacement": (_i, [_dp, _dp, _dp, _dp, _i64, _dp, _vp]),
    "glb_galaxy_shear": (_i, [_i64, _dp, _dp, _dp, _dp, _i64, _dp, _dp, _dp, _i, _dp, _vp]),
    "glb_ellipticity": (_i, [_i, C.c_double, _dp, _i64, C.c_uint64, C.c_uint32, C.c_uint64, _dp, _vp]),
    "glb_gaussian_phz": (_i, [_dp, _dp, C.c_double, _dp, C.c_double, _dp, C.c_double, _dp, _i, _i64, C.c_uint64, C.c_uint32, _dp, _dp, _vp]),
    "glb_redshifts_from_cdf": (_i, [_dp, _dp, _i, _dp, _i64, C.c_uint64, C.c_uint32, C.c_uint64, _dp, _vp]),
    "glb_query_strip": (_i, [_i64, C.c_double, C.c_double, _dp, _vp]),
    "glb_rotate_map_pixel": (_i, [_i64, C.POINTER(C.c_double), _dp, _dp, _vp]),
    "glb_cls_window": (_i, [_i, _i, _i64, _i64, _dp, _dp, _dp, _vp]),
    "glb_effective_cls": (_i, [_i, _i, _i, _i, _i64, _i, _dp, _dp, _dp, _dp, _vp]),
    "glb_alm2map_host": (_i, [_vp, _dp, _i, _dp, _ip, _dp, _vp]),
    "glb_dist_setup": (_i, [_vp, _i, _i, _ip, _ip, _i]),
    "glb_dist_alm2phase": (_i, [_vp, _dp, _i, _dp, _vp]),
    "glb_dist_phase2map": (_i, [_vp, _dp, _i, _dp, _ip, _dp, _vp]),
    "glb_dist_p2p_alloc": (_i, [_vp, _i, _vp]),
    "glb_dist_p2p_open": (_i, [_vp, _vp, _ip]),
    "glb_dist_alm2phase_p2p": (_i, [_vp, _dp, _i, _i, _vp]),
    "glb_dist_p2p_recv": (_i, [_vp, _i, C.POINTER(_vp)]),
    "glb_plan_timing_enable": (_i, [_vp, _i]),
    "glb_plan_timing_read": (_i, [_vp, _dp, _vp, _vp]),
    "glb_kernel_launch_count": (C.c_uint64, []),
    "glb_measure_fp64_peak": (_i, [_i, _dp, _dp, _vp]),
    "glb_debug_alm2phase": (_i, [_vp, _dp, _i, _dp, _vp]),
    "glb_plan_set_legendre_mode": (_i, [_vp, _i]),
    "glb_plan_release_scratch": (_i, [_vp]),
    "glb_alm2map_prepare": (_i, [_vp, _dp, _i, _i, _vp]),
    "glb_alm2map_finish": (_i, [_vp, _i, _i, _dp, _ip, _dp, _vp]),
    "glb_debug_alm2phase_int8": (_i, [_vp, _dp, _i, _dp, _vp]),
    "glb_debug_phase2map": (_i, [_vp, _dp, _i, _dp, _vp]),
    "glb_debug_mlim": (_i, [_vp, _ip]),
}


class GlassB200Error(RuntimeError):
    pass


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load libglassb200.so; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise GlassB200Error(
            f"{_LIB_PATH} not found: build the CUDA extension first "
            "(python -m glass_b200.build, or __graft_entry__.build()); there is no CPU fallback"
        )
    lib = C.CDLL(os.fspath(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    """Map a glb_status to the exception the reference raises for that condition."""
    if status == GLB_OK:
        return
    lib = load()
    detail = (lib.glb_last_error() or b"").decode()
    name = (lib.glb_status_string(status) or b"").decode()
    if status in (GLB_ERR_NOT_POSDEF, GLB_ERR_NEGATIVE_CL):
        raise ValueError(name)
    if status == GLB_ERR_INVALID_ARG:
        raise ValueError(f"{what}: {detail or name}")
    if status == GLB_ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {detail or name}")
    if status == GLB_ERR_NOMEM:
        raise MemoryError(f"{what}: {detail or name}")
    raise GlassB200Error(f"{what}: {name}: {detail}")
