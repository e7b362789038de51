"""ctypes binding of ``libqldpc_b200.so`` (C ABI declared in ``include/qldpc_b200.h``).

There is no CPU fallback: if the shared library is missing this module raises at first use, and
every compute entry point of the library fails with ``QD_ENODEV`` when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libqldpc_b200.so")

QD_BP_PRODUCT_SUM = 0
QD_BP_MIN_SUM = 3
QD_PREC_F64 = 64
QD_PREC_F32 = 32

QD_EINVAL, QD_ENODEV, QD_ECUDA, QD_ENOMEM, QD_EUNSUP = -1, -2, -3, -4, -5

i32, i64, u32, u64, u8, f64 = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_uint8, C.c_double
P = C.POINTER
vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/qldpc_b200.h declares
SIGNATURES = {
    "qd_abi_version": (C.c_int, []),
    "qd_last_error": (C.c_char_p, []),
    "qd_device_count": (C.c_int, []),
    "qd_device_info": (C.c_int, [C.c_int, P(C.c_int), P(C.c_int), P(C.c_int)]),
    "qd_graph_create": (C.c_int, [i32, i32, i64, vp, vp, C.c_int, P(vp)]),
    "qd_graph_destroy": (C.c_int, [vp]),
    "qd_logicals_create": (C.c_int, [i32, i32, vp, C.c_int, P(vp)]),
    "qd_logicals_destroy": (C.c_int, [vp]),
    "qd_sampler_create": (C.c_int, [i32, vp, C.c_int, P(vp)]),
    "qd_sampler_destroy": (C.c_int, [vp]),
    "qd_sampler_run": (C.c_int, [vp, u64, u32, i64, i64, vp, i32, vp]),
    "qd_circuit_create": (C.c_int, [i32, vp, i32, vp, i32, i32, C.c_int, P(vp)]),
    "qd_circuit_destroy": (C.c_int, [vp]),
    "qd_circuit_sample": (C.c_int, [vp, u64, i64, i64, vp, i32, vp]),
    "qd_circuit_propagate_faults": (C.c_int, [vp, i64, vp, vp, i32, vp]),
    "qd_bits_copy_batch": (C.c_int, [vp, i32, i64, i64, i32, vp, i32, i64, i32, vp]),
    "qd_syndrome_batch": (C.c_int, [vp, vp, i32, i64, vp, i32, vp]),
    "qd_logical_fail_batch": (C.c_int, [vp, vp, vp, i32, i64, vp, vp, vp]),
    "qd_bp_create": (C.c_int, [vp, vp, i32, i32, f64, i32, P(vp)]),
    "qd_bp_destroy": (C.c_int, [vp]),
    "qd_bp_configure": (C.c_int, [vp, i32, i32, i32, i32]),
    "qd_bp_get_config": (C.c_int, [vp, P(i32), P(i32), P(i32), P(i32), P(i32), P(i32)]),
    "qd_bp_select_kernel": (C.c_int, [vp, i32]),
    "qd_bp_get_kernel": (C.c_int, [vp, P(i32), P(i32), P(i32)]),
    "qd_bp_cluster_plan_check": (C.c_int, [i32, i32, vp, vp, i32, i32, vp]),
    "qd_bp_decode_batch": (C.c_int, [vp, vp, i32, i64, vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp]),
    "qd_bp_decode_batch_host": (C.c_int, [vp, vp, i32, i64, vp, i32, vp, vp]),
    "qd_bp_decode_one": (C.c_int, [vp, vp, vp, P(i32), P(i32), vp]),
    "qd_osd_create": (C.c_int, [vp, i32, i32, P(vp)]),
    "qd_osd_destroy": (C.c_int, [vp]),
    "qd_osd_rank": (C.c_int, [vp]),
    "qd_bposd_decode_batch": (C.c_int, [vp, vp, vp, i32, i64, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp]),
    "qd_bposd_decode_batch_host": (C.c_int, [vp, vp, vp, i32, i64, vp, i32, vp, vp]),
    "qd_ssf_create": (C.c_int, [vp, i32, vp, vp, i32, P(vp)]),
    "qd_ssf_destroy": (C.c_int, [vp]),
    "qd_ssf_decode_batch": (C.c_int, [vp, vp, i32, i64, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp]),
    "qd_ssf_decode_batch_host": (C.c_int, [vp, vp, i32, i64, vp, i32, vp, vp]),
    "qd_code_capacity_run": (C.c_int, [vp, vp, vp, vp, vp, vp, u64, i64, i64, vp, vp, i32, vp, vp, vp]),
}

_lib = None


class QdError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libqldpc_b200 error {code}: {message}")
        self.code = code


def lib() -> C.CDLL:
    """Load the CUDA extension; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"CUDA extension not built: {LIB_PATH} is missing. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "exp_ldpc_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.qd_abi_version() != 2:
            raise RuntimeError("libqldpc_b200.so ABI version mismatch; rebuild the extension")
        _lib = L
    return _lib


def check(rc: int) -> None:
    """Map a C return code onto the reference's Python exception conventions
    (ValueError for bad shapes/arguments, RuntimeError otherwise; SURVEY.md §8b 'Errors')."""
    if rc == 0:
        return
    msg = lib().qd_last_error().decode("utf-8", "replace")
    if rc == QD_EINVAL:
        raise ValueError(msg)
    if rc == QD_ENOMEM:
        raise MemoryError(msg)
    raise QdError(rc, msg)


def ptr(a):
    """Raw pointer of a numpy array / torch tensor / None."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data
