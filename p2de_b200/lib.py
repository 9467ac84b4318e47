"""ctypes binding of libp2de_b200.so — the same entry points Julia binds with `ccall`.

Fails loudly when the CUDA extension is missing or no GPU is present: there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

from .abi import BCDataC, Config, GeometryC, OperatorsC

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("P2DE_B200_LIB", os.path.join(HERE, "libp2de_b200.so"))

EXPORTS = [
    "p2de_create", "p2de_destroy", "p2de_last_error", "p2de_set_stream", "p2de_set_state", "p2de_get_state",
    "p2de_set_state_async", "p2de_get_state_async", "p2de_synchronize", "p2de_rhs", "p2de_get_field",
    "p2de_ssp33_step", "p2de_ssp33_step_async", "p2de_last_dt", "p2de_ssp33_run", "p2de_reduce",
    "p2de_comm_unique_id", "p2de_comm_init", "p2de_kernel_launch_count", "p2de_device_state_ptr",
    "p2de_profile", "p2de_profile_get", "p2de_debug_counters", "p2de_calculate_error",
    "p2de_snapshot_ring", "p2de_snapshot_count", "p2de_snapshot_get",
]

_LIB = None


class P2DEError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libp2de_b200 error {code}: {msg}")
        self.code = code


def load():
    """dlopen the in-tree library; raise if it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: run `python -m p2de_b200.build` (or __graft_entry__.build()). "
                          "p2de_b200 has no CPU fallback.")
    L = C.CDLL(SO_PATH)
    vp, dp, i32, i64, dbl = C.c_void_p, C.POINTER(C.c_double), C.c_int32, C.c_int64, C.c_double
    L.p2de_create.restype = i32
    L.p2de_create.argtypes = [C.POINTER(Config), C.POINTER(OperatorsC), C.POINTER(GeometryC), C.POINTER(BCDataC), C.POINTER(vp)]
    L.p2de_destroy.restype = i32; L.p2de_destroy.argtypes = [vp]
    L.p2de_last_error.restype = C.c_char_p; L.p2de_last_error.argtypes = [vp]
    L.p2de_set_stream.restype = i32; L.p2de_set_stream.argtypes = [vp, vp]
    for f in ("p2de_set_state", "p2de_get_state", "p2de_set_state_async", "p2de_get_state_async"):
        getattr(L, f).restype = i32; getattr(L, f).argtypes = [vp, vp]
    L.p2de_synchronize.restype = i32; L.p2de_synchronize.argtypes = [vp]
    L.p2de_rhs.restype = i32; L.p2de_rhs.argtypes = [vp, dbl, dbl, i32, dp]
    L.p2de_get_field.restype = i32; L.p2de_get_field.argtypes = [vp, i32, vp, i64]
    L.p2de_ssp33_step.restype = i32; L.p2de_ssp33_step.argtypes = [vp, dbl, dp]
    L.p2de_ssp33_step_async.restype = i32; L.p2de_ssp33_step_async.argtypes = [vp, dbl]
    L.p2de_last_dt.restype = i32; L.p2de_last_dt.argtypes = [vp, dp]
    L.p2de_ssp33_run.restype = i32; L.p2de_ssp33_run.argtypes = [vp, dp, i64, C.POINTER(i64), vp]
    L.p2de_reduce.restype = i32; L.p2de_reduce.argtypes = [vp, i32, dp]
    L.p2de_comm_unique_id.restype = i32; L.p2de_comm_unique_id.argtypes = [vp]
    L.p2de_comm_init.restype = i32; L.p2de_comm_init.argtypes = [vp, i32, i32, vp]
    L.p2de_profile.restype = i32; L.p2de_profile.argtypes = [vp, i32]
    L.p2de_profile_get.restype = i32; L.p2de_profile_get.argtypes = [vp, i32, dp, C.POINTER(i64)]
    L.p2de_calculate_error.restype = i32; L.p2de_calculate_error.argtypes = [vp, vp, vp]
    L.p2de_snapshot_ring.restype = i32; L.p2de_snapshot_ring.argtypes = [vp, i32, i64]
    L.p2de_snapshot_count.restype = i64; L.p2de_snapshot_count.argtypes = [vp]
    L.p2de_snapshot_get.restype = i32; L.p2de_snapshot_get.argtypes = [vp, i64, vp, dp, C.POINTER(i64)]
    L.p2de_debug_counters.restype = i32; L.p2de_debug_counters.argtypes = [vp, i32, vp]
    L.p2de_kernel_launch_count.restype = i64; L.p2de_kernel_launch_count.argtypes = [vp]
    L.p2de_device_state_ptr.restype = vp; L.p2de_device_state_ptr.argtypes = [vp]
    _LIB = L
    return L


def check(L, handle, rc):
    if rc != 0:
        msg = L.p2de_last_error(handle)
        raise P2DEError(rc, msg.decode() if msg else "?")
