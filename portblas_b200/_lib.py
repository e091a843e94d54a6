"""ctypes binding of the C-ABI declared in include/pbx_gemm.h.

Loads portblas_b200/libpbx_gemm.so (built by ``python -m portblas_b200.build``).  There is no
fallback: a missing library or a missing sm_100 device raises.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char, c_char_p, c_int, c_int64, c_void_p, POINTER
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libpbx_gemm.so"

# pbx_dtype_t
F32, F64, F16, F16_F32, BF16, BF16_F32 = range(6)
# pbx_kernel_t
KERNEL_AUTO, KERNEL_SIMT, KERNEL_TCGEN05, KERNEL_DMMA, KERNEL_INTERLEAVED, KERNEL_SCAL, KERNEL_NONE = range(7)
KERNEL_NAMES = ["auto", "simt", "tcgen05", "dmma", "interleaved", "scal", "none"]
# pbx_status_t
OK = 0

# every symbol include/pbx_gemm.h declares: name -> (restype, argtypes)
_GEMM_TAIL = [c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64,
              c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int]
SYMBOLS = {
    "pbx_create": (c_int, [POINTER(c_void_p), c_int, c_void_p]),
    "pbx_destroy": (c_int, [c_void_p]),
    "pbx_set_stream": (c_int, [c_void_p, c_void_p]),
    "pbx_get_stream": (c_void_p, [c_void_p]),
    "pbx_get_num_compute_units": (c_int, [c_void_p]),
    "pbx_get_device": (c_int, [c_void_p]),
    "pbx_synchronize": (c_int, [c_void_p]),
    "pbx_last_error": (c_char_p, [c_void_p]),
    "pbx_status_string": (c_char_p, [c_int]),
    "pbx_reload_env": (c_int, [c_void_p]),
    "pbx_set_forced_kernel": (c_int, [c_void_p, c_int]),
    "pbx_set_split_k": (c_int, [c_void_p, c_int]),
    "pbx_last_kernel": (c_int, [c_void_p]),
    "pbx_last_split_k": (c_int, [c_void_p]),
    "pbx_last_repack": (c_int, [c_void_p]),
    "pbx_last_presplit": (c_int, [c_void_p]),
    "pbx_launch_count": (c_int64, [c_void_p]),
    "pbx_workspace_bytes": (c_int64, [c_void_p]),
    "pbx_plan_query": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, c_int64, POINTER(c_int), POINTER(c_int),
                               POINTER(c_int), POINTER(c_int)]),
    "pbx_gemm": (c_int, [c_void_p, c_int, c_char, c_char] + _GEMM_TAIL),
    "pbx_sgemm": (c_int, [c_void_p, c_char, c_char] + _GEMM_TAIL),
    "pbx_dgemm": (c_int, [c_void_p, c_char, c_char] + _GEMM_TAIL),
    "pbx_hgemm": (c_int, [c_void_p, c_char, c_char] + _GEMM_TAIL),
    "pbx_hsgemm": (c_int, [c_void_p, c_char, c_char] + _GEMM_TAIL),
    "pbx_bf16gemm": (c_int, [c_void_p, c_char, c_char] + _GEMM_TAIL),
    "pbx_scal_matrix": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64]),
    "pbx_symm": (c_int, [c_void_p, c_int, c_char, c_char, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                         c_void_p, c_void_p, c_int64]),
    "pbx_trsm": (c_int, [c_void_p, c_int, c_char, c_char, c_char, c_char, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                         c_void_p, c_int64]),
    "pbx_cgemm": (c_int, [c_void_p, c_char, c_char, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                          c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64]),
    "pbx_zgemm": (c_int, [c_void_p, c_char, c_char, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                          c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64]),
    "pbx_set_conj_transpose": (c_int, [c_void_p, c_int]),
    "pbx_gemm_multicast": (c_int, [c_void_p, c_int, c_char, c_char, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                                   c_void_p, c_int64, c_void_p, POINTER(c_void_p), c_int, c_int64]),
    "pbx_ipc_export": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_int64)]),
    "pbx_ipc_import": (c_int, [c_void_p, c_void_p, c_int64, POINTER(c_void_p)]),
    "pbx_gemm_host": (c_int, [c_void_p, c_int, c_char, c_char] + _GEMM_TAIL),
    "pbx_shard_range": (c_int, [c_int64, c_int, c_int, c_int64, POINTER(c_int64), POINTER(c_int64)]),
    "pbx_multi_create": (c_int, [POINTER(c_void_p), c_int, POINTER(c_int)]),
    "pbx_multi_destroy": (c_int, [c_void_p]),
    "pbx_multi_device_count": (c_int, [c_void_p]),
    "pbx_multi_handle": (c_void_p, [c_void_p, c_int]),
    "pbx_multi_synchronize": (c_int, [c_void_p]),
    "pbx_multi_last_error": (c_char_p, [c_void_p]),
    "pbx_gemm_sharded": (c_int, [c_void_p, c_int, c_char, c_char, c_int64, c_int64, c_int64, c_void_p, POINTER(c_void_p), c_int64,
                                 POINTER(c_void_p), c_int64, c_void_p, POINTER(c_void_p), c_int64, c_int]),
    "pbx_gemm_strided_batched_sharded": (c_int, [c_void_p, c_int, c_char, c_char, c_int64, c_int64, c_int64, c_void_p,
                                                 POINTER(c_void_p), c_int64, c_int64, POINTER(c_void_p), c_int64, c_int64,
                                                 c_void_p, POINTER(c_void_p), c_int64, c_int64, c_int64]),
    "pbx_gemm_sharded_host": (c_int, [c_void_p, c_int, c_char, c_char, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                                      c_void_p, c_int64, c_void_p, c_void_p, c_int64]),
    "pbx_malloc": (c_int, [c_void_p, POINTER(c_void_p), c_int64]),
    "pbx_free": (c_int, [c_void_p, c_void_p]),
    "pbx_copy_to_device": (c_int, [c_void_p, c_void_p, c_void_p, c_int64]),
    "pbx_copy_to_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int64]),
    "pbx_copy2d_to_device": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int]),
    "pbx_copy2d_to_host": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int]),
    "pbx_fill_bytes": (c_int, [c_void_p, c_void_p, c_int, c_int64]),
    "pbx_copy_device_to_device": (c_int, [c_void_p, c_void_p, c_void_p, c_int64]),
    "pbx_fill": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64]),
    "pbx_event_create": (c_int, [c_void_p, POINTER(c_void_p)]),
    "pbx_event_record": (c_int, [c_void_p, c_void_p]),
    "pbx_event_synchronize": (c_int, [c_void_p, c_void_p]),
    "pbx_event_elapsed_ms": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(ctypes.c_float)]),
    "pbx_event_destroy": (c_int, [c_void_p, c_void_p]),
    "pbx_stream_wait_event": (c_int, [c_void_p, c_void_p]),
    "pbx_device_name": (c_int, [c_void_p, c_char_p, c_int]),
}

_lib = None


def load() -> ctypes.CDLL:
    """dlopen the library and type every entry point.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m portblas_b200.build` "
                "(there is no CPU fallback for the GEMM path)")
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
