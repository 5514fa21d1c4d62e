"""ctypes binding of libtimet_b200.so (include/timet_b200.h).  Plumbing only.

There is no CPU fallback anywhere in this package: if the library cannot be loaded (or built
with nvcc when missing/stale) importing the binding raises, and every compute entry point
returns an error code without a CUDA device, which `check` turns into RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

_lib = None


class SinkhornOpts(C.Structure):
    """struct timet_sinkhorn_opts"""
    _fields_ = [("out_block_rows", C.c_int64), ("out_block_stride", C.c_int64), ("share_sm", C.c_int32), ("reserved", C.c_int32)]


class FFParams(C.Structure):
    """struct timet_ff_params"""
    _fields_ = [("n_clips", C.c_int32), ("n_frames", C.c_int32), ("grid_h", C.c_int32), ("grid_w", C.c_int32),
                ("dim", C.c_int32), ("n_channels", C.c_int32), ("n_last_frames", C.c_int32), ("radius", C.c_int32),
                ("topk", C.c_int32), ("t_begin", C.c_int32), ("temperature", C.c_float), ("reserved", C.c_int32)]


# name -> (restype, argtypes); mirrors include/timet_b200.h one to one
_P = C.c_void_p
_PROTOS = {
    "timet_last_error": (C.c_char_p, []),
    "timet_abi_version": (C.c_int, []),
    "timet_launch_count": (C.c_int64, []),
    "timet_debug_reload_env": (C.c_int, []),
    "timet_sinkhorn_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "timet_sinkhorn": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "timet_sinkhorn_ex": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, _P, _P, _P, _P, C.c_size_t, _P]),
    "timet_sinkhorn_pair": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "timet_sinkhorn_pair_mode": (C.c_int, [C.c_int64, C.c_int]),
    "timet_sinkhorn_resident": (C.c_int, [C.c_int64, C.c_int]),
    "timet_cosine_scores_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int]),
    "timet_cosine_scores": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, _P, _P, C.c_size_t, _P]),
    "timet_cosine_scores_multi": (C.c_int, [_P, C.c_int, C.c_int64, _P, C.c_int, C.c_int, _P, _P, C.c_size_t, _P]),
    "timet_ff_workspace_bytes": (C.c_size_t, [C.POINTER(FFParams)]),
    "timet_ff_tc_supported": (C.c_int, [C.POINTER(FFParams)]),
    "timet_ff_tc_executed_flops": (C.c_double, [C.POINTER(FFParams)]),
    "timet_ff_tc_plan": (C.c_int, [C.POINTER(FFParams), _P]),
    "timet_ff_prepare": (C.c_int, [C.POINTER(FFParams), _P, _P, C.c_size_t, _P]),
    "timet_ff_select": (C.c_int, [C.POINTER(FFParams), C.c_int, _P, _P, C.c_size_t, _P]),
    "timet_ff_select_timed": (C.c_int, [C.POINTER(FFParams), C.c_int, _P, _P, C.c_size_t, _P, _P, _P]),
    "timet_ff_gather": (C.c_int, [C.POINTER(FFParams), _P, _P, _P, C.c_size_t, _P]),
    "timet_ff_propagate": (C.c_int, [C.POINTER(FFParams), C.c_int, _P, _P, _P, _P, C.c_size_t, _P]),
    "timet_ff_stats": (C.c_int, [C.POINTER(FFParams), _P, C.c_size_t, _P, _P]),
    "timet_ff_slots": (C.c_int, [C.POINTER(FFParams)]),
    "timet_ff_export_selection": (C.c_int, [C.POINTER(FFParams), _P, C.c_size_t, C.c_int, C.c_int, _P, _P, _P, _P]),
    "timet_ff_export_wide": (C.c_int, [C.POINTER(FFParams), _P, C.c_size_t, C.c_int64, C.c_int, _P, _P, _P]),
    "timet_debug_tc_tile": (C.c_int, [C.POINTER(FFParams), _P, C.c_size_t, C.c_int64, _P, _P]),
    "timet_debug_tc_trace": (C.c_int, [C.POINTER(FFParams), _P, C.c_size_t, _P, C.c_int, _P]),
    "timet_restrict_neighborhood": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P]),
    "timet_norm_mask": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int, _P]),
    "timet_upsample_argmax": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, _P, _P]),
    "timet_comm_unique_id": (C.c_int, [_P]),
    "timet_comm_init": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P)]),
    "timet_comm_p2p_handle": (C.c_int, [_P, _P]),
    "timet_comm_p2p_connect": (C.c_int, [_P, _P]),
    "timet_comm_p2p_disable": (C.c_int, [_P]),
    "timet_comm_destroy": (C.c_int, [_P]),
    "timet_comm_allreduce_f32": (C.c_int, [_P, _P, C.c_int64, _P]),
}
EXPORTS = tuple(_PROTOS)

SK_EXP, SK_SCORES = 0, 1
FF_EXACT, FF_TC, FF_AUTO = 0, 1, 2
UNIQUE_ID_BYTES = 128
IPC_HANDLE_BYTES = 64


def lib():
    global _lib
    if _lib is None:
        if not _build.is_current():
            if os.environ.get("TIMET_NO_BUILD") == "1" and os.path.isfile(_build.LIB):
                pass                      # use the shipped binary as is
            else:
                _build.build()            # raises if nvcc is missing: no fallback
        handle = C.CDLL(_build.LIB)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(handle, name)    # AttributeError if the ABI lost a symbol
            fn.restype, fn.argtypes = res, args
        if handle.timet_abi_version() != 2:
            raise RuntimeError("libtimet_b200 ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().timet_last_error().decode(errors="replace")
        raise RuntimeError(f"libtimet_b200 {what} failed ({rc}): {msg}")


def reload_env() -> None:
    """Re-read the TIMET_* experiment switches (the library reads the environment once, at first use)."""
    lib().timet_debug_reload_env()


def launch_count() -> int:
    return int(lib().timet_launch_count())
