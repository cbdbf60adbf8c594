"""ctypes binding of the C-ABI in ``include/ms_b200.h`` (csrc/libms_b200.so).

There is no CPU fallback: if the shared library is missing or a call fails the
caller gets an exception, never a silently different code path.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

MS_OK = 0
MS_LIVE_HIST_MAX = 256
_ERR_NAMES = {-1: "MS_ERR_INVALID_ARG", -2: "MS_ERR_UNSUPPORTED", -3: "MS_ERR_CUDA", -4: "MS_ERR_WORKSPACE"}


class MsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class MsUnsupported(MsError):
    """The requested geometry is outside what a specialised kernel supports."""


class LiveConfig(C.Structure):
    _fields_ = [("block_samples", C.c_int64), ("fs", C.c_double), ("k_std", C.c_double),
                ("init_wait_sec", C.c_double), ("after_wait_sec", C.c_double), ("mean_min_db", C.c_double),
                ("dur_min_sec", C.c_double), ("avg_win", C.c_int32), ("reserved", C.c_int32)]


class LiveState(C.Structure):
    _fields_ = [("block_index", C.c_int64), ("state", C.c_int32), ("hist_len", C.c_int32),
                ("hist_pos", C.c_int32), ("trk_n", C.c_int32), ("locked_threshold", C.c_double),
                ("lock_until_sec", C.c_double), ("trk_t0", C.c_double), ("trk_sum", C.c_double),
                ("trk_min", C.c_double), ("trk_max", C.c_double), ("trk_mean_run", C.c_double),
                ("trk_m2_run", C.c_double), ("hist", C.c_double * MS_LIVE_HIST_MAX)]


_p = C.c_void_p
_i32, _i64, _f64 = C.c_int32, C.c_int64, C.c_double

# name -> (restype, argtypes); mirrors include/ms_b200.h declaration by declaration
SIGNATURES = {
    "ms_abi_version": (C.c_int, []),
    "ms_last_error": (C.c_char_p, []),
    "ms_band_power_i16": (C.c_int, [_p, _i64, _i64, _i64, _i32, _i32, _p, _i32, _i32, _i32, _i32, _i32, _i64,
                                    _p, _p, _p, _p, _p]),
    "ms_band_power_f32": (C.c_int, [_p, _i64, _i64, _i64, _i32, _i32, _p, _i32, _i32, _i32, _i32, _i32, _i64,
                                    _p, _p, _p, _p, _p]),
    "ms_dft_i8_plan_bytes": (_i64, [_i32, _i32]),
    "ms_dft_i8_plan_build": (C.c_int, [_p, _p, _i32, _i32, _p, _p]),
    "ms_band_power_i16_tc": (C.c_int, [_p, _i64, _i64, _p, _i32, _i32, _p, _p, _p, _p, _p]),
    "ms_band_power_i16_tc_batched": (C.c_int, [_p, _i64, _i64, _i64, _i64, _p, _i32, _i32, _i64, _p, _p, _p, _p, _p]),
    "ms_dft_seg_plan_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "ms_dft_seg_plan_build": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _p, _p]),
    "ms_band_power_i16_seg": (C.c_int, [_p, _i64, _i64, _i64, _i64, _i64, _p, _i32, _i32, _i32, _i32, _i64, _i64,
                                        _p, _p, _p, _p, _i32, _i32, _p]),
    "ms_dft_seg_projections_i16": (C.c_int, [_p, _i64, _i64, _i64, _i64, _p, _i32, _i32, _i64, _p, _i32, _i32, _p]),
    "ms_window_combine": (C.c_int, [_p, _p, _i64, _i64, _i64, _i32, _i32, _i32, _p, _i32, _i32, _i32, _i32, _i64,
                                    _p, _p, _p, _p, _p]),
    "ms_detect_workspace_bytes": (_i64, [_i64, _i64]),
    "ms_detect_global": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _f64, _i32, _p, _p, _p, _p, _p, _f64, _p, _i64, _p]),
    "ms_detect_adaptive": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _f64, _i32, _i32, _i32, _i32, _i32, _p, _p, _p,
                                     _p, _p, _f64, _p, _i64, _p]),
    "ms_detect_adaptive_hourly": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _f64, _i32, _i32, _i32, _i32, _i32, _p, _p,
                                            _p, _p, _p, _f64, _p, _i64, _p, _f64, _f64, _i64, _i32, _p, C.c_uint32,
                                            _p]),
    "ms_hourly_counts": (C.c_int, [_p, _p, _i64, _i32, _p, _f64, _f64, _i64, _i32, _p, _p]),
    "ms_detector_a_pass_i16": (C.c_int, [_p, _i64, _i64, _i32, _p, _i32, _i32, _f64, _i32, _i32, _i32, _i32, _i32,
                                         _p, _p, _p, _p, _p, _p, _i64, _p, _f64, _f64, _i64, _i32, _p, _p, _p, _p]),
    "ms_detector_a_pass_overlapped_i16": (C.c_int, [_p, _i64, _i64, _i32, _p, _i32, _i32, _f64, _i32, _i32, _i32, _i32,
                                                    _i32, _p, _p, _p, _p, _p, _p, _i64, _p, _f64, _f64, _i64, _i32, _p,
                                                    _p, _p, _p, _p, _p, _p]),
    "ms_read_files": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32]),
    "ms_ingest_rows_h2d": (C.c_int, [_p, _i64, _i64, _i64, _p, _i64, _p]),
    "ms_welch_band_db_f32": (C.c_int, [_p, _i64, _i64, _i64, _i32, _i32, _p, _i32, _p, _f64, _p, _i32, _i32, _p, _p]),
    "ms_welch_band_db_i16": (C.c_int, [_p, _i64, _i64, _i64, _i32, _i32, _p, _i32, _p, _f64, _p, _i32, _i32, _p, _p]),
    "ms_welch_band_db_qf_i16": (C.c_int, [_p, _i64, _i64, _i64, _i32, _i32, _p, _p, _p, _p]),
    "ms_welch_band_db_qf_f32": (C.c_int, [_p, _i64, _i64, _i64, _i32, _i32, _p, _p, _p, _p]),
    "ms_welch_i8_plan_bytes": (_i64, [_i32]),
    "ms_welch_i8_plan_build": (C.c_int, [_p, _i32, _i32, _p, _p]),
    "ms_welch_band_db_i8_i16": (C.c_int, [_p, _i64, _i64, _i64, _i32, _i32, _p, _p, _p, _p]),
    "ms_live_state_step": (C.c_int, [_p, C.POINTER(LiveConfig), _i64, _p, _i64, _i32, _i64, _i32, _p, _p, _p, _p]),
    "ms_live_state_workspace_bytes": (_i64, [_i64, _i64]),
    "ms_live_state_step_ws": (C.c_int, [_p, C.POINTER(LiveConfig), _i64, _p, _i64, _i32, _i64, _i32, _p, _p, _p, _p,
                                        _i64, _p]),
    "ms_psd_spectrogram_i16": (C.c_int, [_p, _i64, _i64, _i64, _i32, _i32, _p, _f64, _i32, _i32, _i32, _i32,
                                         _p, _p, _p]),
    "ms_psd_spectrogram_f32": (C.c_int, [_p, _i64, _i64, _i64, _i32, _i32, _p, _f64, _i32, _i32, _i32, _i32,
                                         _p, _p, _p]),
}

_lib = None


def load() -> C.CDLL:
    """Load csrc/libms_b200.so; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `python -m meteor_scatter_b200.build`). "
            "There is no CPU fallback for the detection path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.ms_abi_version() != 1:
        raise RuntimeError(f"libms_b200.so ABI version {lib.ms_abi_version()} != 1; rebuild")
    _lib = lib
    return lib


def check(code: int):
    if code != MS_OK:
        msg = load().ms_last_error().decode("utf-8", "replace")
        raise (MsUnsupported if code == -2 else MsError)(code, msg)


def ptr(t):
    """Device/host pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
