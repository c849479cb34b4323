"""ctypes binding of libsglb200.so (the C-ABI CUDA library, include/sglb200.h).

The product path has NO CPU fallback: if the shared object is missing or no B200 is visible, every call raises
SglB200Error.  Build the library with ``python sgl_b200/csrc/build.py`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libsglb200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_ALLOC = 0, 1, 2, 3, 4
HOST, DEVICE = 0, 1
MODE_FAST, MODE_EXACT = 0, 1
AGG_SUM, AGG_MEAN, AGG_MAX, AGG_MIN, AGG_WEIGHTED, AGG_CONCAT, AGG_OSD, AGG_LAST = range(8)
LW_KINDS = {"simple": 0, "simple_allow_neg": 1, "gate": 2, "ori_ref": 3, "jk": 4}

# every symbol include/sglb200.h declares (tests/test_abi.py checks the header against this table and the .so)
SIGNATURES = {
    "sglb200_version": (c_int, []),
    "sglb200_last_error": (c_char_p, []),
    "sglb200_device_count": (c_int, []),
    "sglb200_set_device": (c_int, [c_int]),
    "sglb200_graph_create": (c_int, [POINTER(c_void_p), c_int64, c_int64, c_int64, c_void_p, c_int, c_void_p, c_void_p,
                                     c_int, c_int, c_int, c_void_p]),
    "sglb200_graph_destroy": (c_int, [c_void_p]),
    "sglb200_graph_set_values": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "sglb200_graph_info": (c_int, [c_void_p, POINTER(c_int64)]),
    "sglb200_normalize_values": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_int, c_int, c_void_p]),
    "sglb200_adjacency_build": (c_int, [POINTER(c_void_p), c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int,
                                        POINTER(c_int64), c_void_p]),
    "sglb200_adjacency_export": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sglb200_adjacency_free": (None, [c_void_p]),
    "sglb200_spmm": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "sglb200_graph_chunks": (c_int, [c_void_p, c_int, c_int, POINTER(c_int64), POINTER(c_int64)]),
    "sglb200_spmm_tiles": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int64, c_int64,
                                   c_void_p]),
    "sglb200_propagate": (c_int, [c_void_p, POINTER(c_void_p), c_int64, c_int, c_int, c_int, c_void_p]),
    "sglb200_propagate_fused": (c_int, [c_void_p, c_void_p, c_int64, POINTER(c_void_p), c_int64, c_int, c_int, c_int, c_int,
                                        c_int, c_int, POINTER(c_float), c_void_p, c_int64, c_int, c_void_p]),
    "sglb200_spmm_axpby": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_int64,
                                   c_int, c_float, c_float, c_void_p]),
    "sglb200_propagate_host": (c_int, [c_void_p, c_void_p, POINTER(c_void_p), c_int, c_int, c_int]),
    "sglb200_aggregate": (c_int, [c_int, POINTER(c_void_p), c_int, c_int64, c_int, c_int64, POINTER(c_float), c_void_p,
                                  c_int64, c_void_p]),
    "sglb200_lw_forward": (c_int, [c_int, POINTER(c_void_p), c_int, c_int, c_int, c_int64, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "sglb200_lw_backward": (c_int, [c_int, POINTER(c_void_p), c_int, c_int, c_int, c_int64, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, POINTER(c_void_p), c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "sglb200_it_forward": (c_int, [POINTER(c_void_p), c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p]),
    "sglb200_it_backward": (c_int, [POINTER(c_void_p), c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                    POINTER(c_void_p), c_void_p, c_void_p, c_void_p]),
    "sglb200_relu_concat": (c_int, [POINTER(c_void_p), c_int, c_int64, c_int, c_void_p, c_void_p]),
    "sglb200_relu_concat_backward": (c_int, [POINTER(c_void_p), c_int, c_int64, c_int, c_void_p, POINTER(c_void_p), c_void_p]),
    "sglb200_gather_rows": (c_int, [POINTER(c_void_p), c_int, c_int64, c_void_p, c_int64, c_int, POINTER(c_void_p),
                                    c_int64, c_void_p]),
    "sglb200_ipc_alloc": (c_int, [c_int64, POINTER(c_void_p), c_void_p]),
    "sglb200_ipc_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "sglb200_ipc_close": (c_int, [c_void_p]),
    "sglb200_ipc_free": (c_int, [c_void_p]),
    "sglb200_push_rows": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p]),
    "sglb200_signal_peers": (c_int, [c_void_p, c_int, ctypes.c_uint64, c_void_p]),
    "sglb200_wait_flags": (c_int, [c_void_p, c_int, ctypes.c_uint64, c_void_p]),
    "sglb200_peer_status": (c_int, []),
    "FloatCSRMulDenseOMP": (None, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "FloatCSRMulDense": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
}


class SglB200Error(RuntimeError):
    """Raised when the CUDA library is missing, no B200 is visible, or a library call fails."""


_lib = None


def load():
    """dlopen libsglb200.so (no CUDA call is made) and attach the prototypes."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SglB200Error(
                f"{LIB_PATH} is missing: build it with `python sgl_b200/csrc/build.py`; sgl_b200 has no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    msg = load().sglb200_last_error()
    return msg.decode() if msg else ""


def check(status: int, what: str = "") -> None:
    if status != OK:
        raise SglB200Error(f"{what or 'libsglb200'} failed (status {status}): {last_error()}")


def ptr_array(ptrs):
    arr = (c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
