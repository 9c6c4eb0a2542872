"""ctypes binding of libmbrf.so — every symbol ``include/mbrf.h`` declares.

The library is the product; this module only loads it.  A missing library is a hard
error (no CPU path exists to fall back to).
"""
from __future__ import annotations

import ctypes as C
import os
import re

PKG = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(PKG, "libmbrf.so")
HEADER = os.path.join(PKG, "..", "include", "mbrf.h")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class MbrfError(RuntimeError):
    """A libmbrf entry point returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libmbrf error {code}: {message}")
        self.code = code
        self.message = message


def library_path() -> str:
    return _LIB_PATH


_dp, _i, _d, _ll, _vp = c_double_p, C.c_int, C.c_double, C.c_longlong, C.c_void_p

# name -> (restype, argtypes); kept in the order of include/mbrf.h
SIGNATURES = {
    "mbrf_version": (C.c_char_p, []),
    "mbrf_last_error": (C.c_char_p, []),
    "mbrf_device_count": (_i, []),
    "mbrf_set_device": (_i, [_i]),
    "mbrf_get_device": (_i, [c_int_p]),
    "mbrf_device_sm_count": (_i, [c_int_p]),
    "mbrf_set_fanout": (_i, [_i]),
    "mbrf_get_fanout": (_i, []),
    "mbrf_peer_alloc": (_i, [C.c_ulonglong, C.POINTER(C.c_void_p), C.c_char_p]),
    "mbrf_peer_open": (_i, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "mbrf_peer_close": (_i, [_vp]),
    "mbrf_peer_free": (_i, [_vp]),
    "mbrf_launch_count": (C.c_ulonglong, []),
    "mbrf_measure_fp64_peak": (_i, [c_double_p, c_double_p]),
    "mbrf_blochsimfz": (_i, [_dp, _dp, _dp, _dp, _dp, _dp, _i, _d, _d, _dp, _i, _dp, _dp, _dp, _i,
                             _dp, _dp, _dp, _i, _d]),
    "mbrf_bloch": (_i, [_dp, _dp, _i, _dp, _i, _dp, _i, _d, _d, _dp, _i, _dp, _i, _i, _i,
                        _dp, _dp, _dp, _i, _dp, _dp, _dp, c_int_p, _d]),
    "mbrf_bloch_workspace_bytes": (C.c_ulonglong, [_i]),
    "mbrf_bloch_device": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _d, _d, _vp, _i, _vp, _vp, _vp, _i,
                               _ll, _ll, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _d, _vp, _vp]),
    "mbrf_bloch_scale_sweep_device": (_i, [_vp, _vp, _vp, _i, _d, _d, _vp, _i, _vp, _i, _ll, _ll,
                                           _vp, _vp, _vp, _d, _vp, _vp]),
    "mbrf_bloch_scale_sweep": (_i, [_dp, _dp, _i, _d, _d, _d, _dp, _i, _dp, _i, _dp, _dp, _dp, _d]),
    "mbrf_bloch_set_tuning": (_i, [_i, _i]),
    "mbrf_abr": (_i, [_dp, _dp, _dp, _dp, _i, _dp, _i, _dp, _i, _i, _dp, _dp, _dp, _dp]),
    "mbrf_abr_device": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _i, _ll, _ll,
                             _vp, _vp, _vp, _vp, _vp, _vp]),
    "mbrf_abr_workspace_bytes": (C.c_ulonglong, [_i]),
    "mbrf_pdhg_padded_sizes": (_i, [_i, _i, _i, c_int_p, c_int_p, c_int_p]),
    "mbrf_pdhg_set_gemm": (_i, [_i]),
    "mbrf_pdhg_set_tc_digits": (_i, [_i]),
    "mbrf_pdhg_set_halpern": (_i, [_i]),
    "mbrf_tc_product_device": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "mbrf_pdhg_set_option": (_i, [_i, _d]),
    "mbrf_b2a_batch": (_i, [_dp, _dp, _i, _i, _dp, _dp]),
    "mbrf_ab2rf_batch": (_i, [_dp, _dp, _dp, _dp, _i, _i, _dp, _dp]),
    "mbrf_fmp2_max_taps": (_i, []),
    "mbrf_fmp2_batch": (_i, [_dp, _dp, _i, _i, _dp, _dp]),
    "mbrf_fmp2_workspace_bytes": (C.c_ulonglong, [_i]),
    "mbrf_fmp2_batch_device": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "mbrf_fir_pdhg_warm_start": (_i, [_dp, _dp, _dp, _dp, _dp]),
    "mbrf_pdhg_warm_start_device": (_i, [_vp, _vp, _vp, _vp]),
    "mbrf_pdhg_workspace_bytes": (C.c_ulonglong, [_i, _i, _i]),
    "mbrf_pdhg_solve_device": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _vp,
                                    _vp, _i, _i, _d, _d, _d, _vp, _vp, _vp, _vp, _vp]),
    "mbrf_fir_pdhg_solve2": (_i, [_dp, _dp, _dp, _i, c_int_p, _dp, _dp, _i, _i, c_int_p, c_int_p, _dp,
                                  c_int_p, c_int_p, _i, _dp, _dp, _dp, _dp, _dp, _dp, _i, _dp, _vp,
                                  _i, _i, _d, _d, _d, _dp, _dp, _dp]),
    "mbrf_fir_ipm_solve": (_i, [_dp, _i, c_int_p, _dp, _dp, _i, c_int_p, c_int_p, _i, _dp, _dp, _dp, _dp, _dp, _dp, _i, _i, _i,
                                _dp, _i, _d, _d, _d, _dp, _dp]),
    "mbrf_fir_ap_solve": (_i, [_i, _i, _dp, _dp, _dp, _dp, _dp, _i, _i, _i, _d, _d, _d, _dp, _dp, _dp, _dp, c_int_p]),
    "mbrf_fir_ap_assemble": (_i, [_i, _i, _dp, _dp, _dp, _dp, _dp, _i, _i, c_int_p, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "mbrf_ipm_padded_sizes": (_i, [_i, _i, _i, c_int_p, c_int_p, c_int_p]),
    "mbrf_ipm_set_option": (_i, [_i, _d]),
    "mbrf_ipm_cholesky_bench": (_i, [_i, _i, _i, _i, C.POINTER(C.c_float)]),
    "mbrf_flip_zero_max_taps": (_i, []),
    "mbrf_flip_zero_batch": (_i, [_dp, _dp, _i, c_int_p, _i, C.POINTER(C.c_ubyte), _i, _d, _d, c_int_p, _dp, _dp, _dp, _dp, _dp,
                                  _dp]),
    "mbrf_fir_pdhg_solve": (_i, [_dp, _dp, _i, c_int_p, _dp, _dp, _i, _i, c_int_p, c_int_p, _i,
                                 _dp, _dp, _dp, _dp, _dp, _dp, _i, _dp, _i, _i, _dp, _i, _i, _d, _d, _d,
                                 _dp, _dp, _dp]),
}


class PdhgBlocks(C.Structure):
    """mbrf_pdhg_blocks (include/mbrf.h)."""
    _fields_ = [("simplex_row0", C.c_int), ("simplex_rows", C.c_int), ("simplex_w", c_double_p),
                ("disk_row0", C.c_int), ("disk_pairs", C.c_int),
                ("group_row0", C.c_int), ("group_pairs", C.c_int), ("group_w", c_double_p),
                ("norm_coords", C.c_int), ("norm_w", c_double_p),
                ("group2_row0", C.c_int), ("group2_pairs", C.c_int), ("group2_w", c_double_p)]


def declared_symbols() -> list[str]:
    """Function names declared in include/mbrf.h (parsed, so the header is the source of truth)."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mbrf_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib() -> C.CDLL:
    """Load libmbrf.so (once).  Raises if it has not been built — there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} is missing: build it with `python -m multiband_rf_pulse_design_b200._build` "
                "(needs nvcc).  libmbrf has no CPU fallback.")
        handle = C.CDLL(_LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise MbrfError(rc, lib().mbrf_last_error().decode(errors="replace"))
