"""ctypes binding of libadtfe.so (include/adtfe.h).  There is no fallback: if the
library cannot be loaded, or the device is not sm_100, every entry raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
#: ``ADTFE_LIB`` selects a tuning variant built by ``build.py --out=...`` (same ABI, same sources)
LIB_PATH = os.environ.get("ADTFE_LIB") or os.path.join(_HERE, "libadtfe.so")

EXPORTS = [
    "adtfe_version", "adtfe_last_error", "adtfe_device_ok",
    "adtfe_bank_create", "adtfe_bank_destroy", "adtfe_bank_bytes",
    "adtfe_render_workspace_bytes", "adtfe_render",
    "adtfe_mel_create", "adtfe_mel_destroy", "adtfe_mel_frames", "adtfe_mel_fast_path", "adtfe_logmel", "adtfe_logmel_rows",
    "adtfe_render_logmel", "adtfe_mel_force_generic", "adtfe_frontend_host", "adtfe_plan_blob_layout",
    "adtfe_resampler_create", "adtfe_resampler_destroy", "adtfe_resample_length", "adtfe_resample", "adtfe_downmix",
    "adtfe_peak_normalise",
    "adtfe_trace_begin", "adtfe_trace_dump",
    "adtfe_planner_create", "adtfe_planner_destroy", "adtfe_planner_plan", "adtfe_planner_export", "adtfe_planner_pack_batches",
    "adtfe_planner_set_fx", "adtfe_planner_export_fx",
    "adtfe_linear_create", "adtfe_linear_destroy", "adtfe_linear_forward",
]


class AdtfeError(RuntimeError):
    pass


class Plan(C.Structure):
    """adtfe_plan"""
    _fields_ = [("events_dev", C.c_void_p), ("segments_dev", C.c_void_p), ("tile_ptr_dev", C.c_void_p),
                ("tile_events_dev", C.c_void_p), ("peak_work_dev", C.c_void_p),
                ("n_events", C.c_int32), ("n_seg", C.c_int32), ("tiles_per_seg", C.c_int32),
                ("n_peak_work", C.c_int32), ("ld_wav", C.c_int64),
                ("mel_rows_dev", C.c_void_p), ("mel_total_rows", C.c_int64), ("mel_max_count", C.c_int32),
                ("n_chunks", C.c_int32), ("chunks_host", C.c_void_p), ("n_tile_events", C.c_int32),
                ("n_fx", C.c_int32), ("fx_dev", C.c_void_p), ("sample_rate", C.c_int32)]


_lock = threading.Lock()
_lib = None


def _declare(lib) -> None:
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    lib.adtfe_version.restype = C.c_int
    lib.adtfe_last_error.restype = C.c_char_p
    lib.adtfe_device_ok.argtypes = [C.c_int]
    lib.adtfe_bank_create.argtypes = [vp, i64, vp, vp, i32, C.c_int, C.POINTER(vp)]
    lib.adtfe_bank_destroy.argtypes = [vp]
    lib.adtfe_bank_bytes.argtypes = [vp]
    lib.adtfe_bank_bytes.restype = i64
    lib.adtfe_render_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.adtfe_render_workspace_bytes.restype = sz
    lib.adtfe_render.argtypes = [vp, C.POINTER(Plan), vp, vp, sz, vp]
    lib.adtfe_mel_force_generic.argtypes = [vp, i32]
    lib.adtfe_mel_create.argtypes = [i32, i32, i32, vp, vp, C.c_int, C.POINTER(vp)]
    lib.adtfe_mel_destroy.argtypes = [vp]
    lib.adtfe_mel_fast_path.argtypes = [vp]
    lib.adtfe_mel_frames.argtypes = [vp, i64, C.POINTER(i32), C.POINTER(i32)]
    lib.adtfe_logmel.argtypes = [vp, vp, i32, i64, i64, vp, vp]
    lib.adtfe_logmel_rows.argtypes = [vp, vp, i32, i64, vp, i32, vp, vp]
    lib.adtfe_render_logmel.argtypes = [vp, vp, C.POINTER(Plan), i64, vp, vp, vp, sz, vp]
    lib.adtfe_frontend_host.argtypes = [vp, vp, C.POINTER(Plan), i64, vp, sz, vp, vp, vp, vp, sz, vp, vp, vp, vp]
    lib.adtfe_plan_blob_layout.argtypes = [C.POINTER(Plan), C.POINTER(sz * 7), C.POINTER(sz)]
    lib.adtfe_resampler_create.argtypes = [i32, i32, i32, vp, C.c_int, C.POINTER(vp)]
    lib.adtfe_resampler_destroy.argtypes = [vp]
    lib.adtfe_resample_length.argtypes = [vp, i64]
    lib.adtfe_resample_length.restype = i64
    lib.adtfe_resample.argtypes = [vp, vp, i32, i64, i64, vp, i64, vp, vp]
    lib.adtfe_downmix.argtypes = [vp, i32, i64, i64, vp, vp]
    lib.adtfe_peak_normalise.argtypes = [vp, i64, vp, i32, vp]
    lib.adtfe_trace_dump.argtypes = [C.c_char_p]
    lib.adtfe_planner_create.argtypes = [i32, C.c_double, C.c_double, C.c_double, i32, vp, vp, i32, vp, vp, vp, vp, vp,
                                         vp, C.POINTER(vp)]
    lib.adtfe_planner_destroy.argtypes = [vp]
    lib.adtfe_planner_plan.argtypes = [vp, vp, vp, i32, vp, i64, vp, vp]
    lib.adtfe_planner_export.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    lib.adtfe_planner_pack_batches.argtypes = [vp, vp, i32, i32, i32, i32, vp, sz, C.POINTER(Plan), vp, vp, vp,
                                               C.POINTER(sz), C.POINTER(sz)]
    lib.adtfe_planner_set_fx.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    lib.adtfe_planner_export_fx.argtypes = [vp, vp]
    lib.adtfe_linear_create.argtypes = [i32, i32, vp, vp, C.c_int, C.POINTER(vp)]
    lib.adtfe_linear_destroy.argtypes = [vp]
    lib.adtfe_linear_forward.argtypes = [vp, vp, i64, vp, vp]


def load():
    """The loaded library.  Raises AdtfeError when libadtfe.so is missing - build it with
    ``python -m adt_str_b200.build`` (or ``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise AdtfeError(f"{LIB_PATH} not found: the CUDA library has not been built "
                                     "(python -m adt_str_b200.build); there is no CPU fallback")
                lib = C.CDLL(LIB_PATH)
                _declare(lib)
                _lib = lib
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().adtfe_last_error().decode(errors="replace")
        raise AdtfeError(f"{what or 'adtfe'} failed ({status}): {msg}")
