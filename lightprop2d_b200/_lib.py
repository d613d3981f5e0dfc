"""ctypes binding of liblp2d.so (C ABI declared in include/lp2d.h).

There is no CPU fallback: if the library is missing it is built with nvcc; if that
is impossible, or no CUDA device is present when a plan is created, the error is
raised to the caller.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblp2d.so")

C64, C128 = 0, 1
MASK_COMPLEX, MASK_REAL, MASK_U8 = 1, 2, 3
AXIS_ROWS, AXIS_COLS = 0, 1
OP_FFT, OP_H_FFT, OP_FFT_H_IFFT = 0, 1, 2
PW_H, PW_LENS = 1, 2

# name -> (restype, argtypes); mirrors include/lp2d.h one to one
_vp, _ll, _i, _d = C.c_void_p, C.c_longlong, C.c_int, C.c_double
_dp = C.POINTER(C.c_double)
SIGNATURES = {
    "lp2d_version": (_i, []),
    "lp2d_last_error": (C.c_char_p, []),
    "lp2d_min_npoints": (_i, []),
    "lp2d_max_npoints": (_i, [_i]),
    "lp2d_max_line_points": (_i, [_i]),
    "lp2d_grid_tables": (_i, [_i, _d, _d, _dp, _dp, _dp, _dp, _dp]),
    "lp2d_plan_create": (_i, [C.POINTER(_vp), _i, _i, _dp, _dp, _dp]),
    "lp2d_plan_destroy": (_i, [_vp]),
    "lp2d_fft2": (_i, [_vp, _vp, _vp, _ll, _i, _vp]),
    "lp2d_propagate": (_i, [_vp, _vp, _vp, _vp, _ll, _d, _vp]),
    "lp2d_propagate_sweep": (_i, [_vp, _vp, _vp, _ll, _i, _d, _ll, _vp, _d, _vp, _vp]),
    "lp2d_propagate_field": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _ll, _d, _vp]),
    "lp2d_lens": (_i, [_vp, _vp, _vp, _ll, _d, _d, _vp]),
    "lp2d_mul_fft2": (_i, [_vp, _vp, _vp, _vp, _i, _ll, _ll, _i, _vp]),
    "lp2d_pass": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _ll, _i, _d, _vp]),
    "lp2d_pass_lines": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _d, _vp]),
    "lp2d_pass_lines_tiles": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _d, _vp]),
    "lp2d_pack_transpose": (_i, [_i, _vp, _vp, _i, _i, _i, _vp]),
    "lp2d_intensity": (_i, [_vp, _vp, _vp, _ll, _vp]),
    "lp2d_expand": (_i, [_i, _vp, _vp, _i, _i, _ll, _vp]),
    "lp2d_coarse": (_i, [_i, _vp, _vp, _i, _i, _ll, _vp]),
    "lp2d_random_fields": (_i, [_i, _vp, _vp, _i, _ll, C.c_ulonglong, C.c_ulonglong, _i, _vp]),
    "lp2d_phase": (_i, [_i, _vp, _vp, _ll, _vp]),
    "lp2d_moments": (_i, [_i, _vp, _i, _d, _d, _d, _vp, _vp]),
    "lp2d_modes_project": (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _ll, _vp]),
    "lp2d_modes_project_flat": (_i, [_i, _vp, _vp, _vp, _ll, _ll, _i, _vp]),
    "lp2d_modes_synthesize": (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _ll, _vp]),
    "lp2d_pointwise": (_i, [_i, _vp, _ll, _i, _i, _i, _i, _vp, _d, _d, _vp]),
    "lp2d_mul_mask": (_i, [_i, _vp, _vp, _i, _ll, _ll, _ll, _vp]),
    "lp2d_intensity_n": (_i, [_i, _vp, _vp, _ll, _vp]),
    "lp2d_bluestein_size": (_i, [_i]),
    "lp2d_bluestein_pre": (_i, [_i, _vp, _vp, _i, _i, _ll, _i, _vp]),
    "lp2d_bluestein_post": (_i, [_i, _vp, _vp, _i, _i, _ll, _i, _vp]),
    "lp2d_bluestein_filter": (_i, [_i, _vp, _i, _i, _i, _vp]),
    "lp2d_solve": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "lp2d_resample": (_i, [_i, _vp, _ll, _i, _vp, _i, _vp, _vp, _vp, _vp, _ll, _vp]),
    "lp2d_plan_step_kernel": (_i, [_vp, C.c_char_p, _i]),
    "lp2d_q4_timing": (_i, [_vp, C.POINTER(C.c_ulonglong), _i]),
    "lp2d_plan_set_spectrum_layout": (_i, [_vp, _i]),
    "lp2d_plan_get_spectrum_layout": (_i, [_vp]),
    "lp2d_transpose": (_i, [_i, _vp, _vp, _i, _ll, _vp]),
    "lp2d_generate": (_i, [_i, _vp, _i, _ll, _vp, _vp, _i, C.POINTER(C.c_double), _i, C.c_ulonglong, C.c_ulonglong, _vp]),
    "lp2d_launch_count": (_ll, []),
}

_lib = None


class Lp2dError(RuntimeError):
    pass


def lib_path():
    return _LIB_PATH


def load():
    """Load (building first if needed) and return the ctypes handle."""
    global _lib
    if _lib is not None:
        return _lib
    # digest-stamped: a no-op when the library matches the sources, a rebuild when they were edited
    from . import build as _build
    try:
        _build.build()
    except RuntimeError as e:
        # no compiler on this machine: a prebuilt library is acceptable, a compile error never is
        if "nvcc not found" not in str(e) or not os.path.exists(_LIB_PATH):
            raise
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().lp2d_last_error().decode("utf-8", "replace")
        if rc == 2:
            raise NotImplementedError(f"lp2d {what}: {msg}")
        if rc == 1:
            raise ValueError(f"lp2d {what}: {msg}")
        raise Lp2dError(f"lp2d {what} failed (code {rc}): {msg}")


def launch_count():
    return int(load().lp2d_launch_count())
