"""ctypes binding of libnepb200.so (the C ABI declared in include/nepb200.h).

This is the Python twin of the Julia `ccall` layer (julia/NEPB200.jl): thin, no arithmetic.  The
library is loaded from the package directory (built in-tree by csrc/Makefile); if it is missing the
import fails loudly -- there is no CPU fallback on the product path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnepb200.so")

NEPB_OK = 0
NEPB_E_INVALID = -1
NEPB_E_CUDA = -2
NEPB_E_SINGULAR = -3
NEPB_E_NOMEM = -4
NEPB_E_UNSUPPORTED = -5

COEF_SCALAR = 0
COEF_DIAG = 1
COEF_GENERAL = 2


class NepbError(RuntimeError):
    """Non-zero status from the C ABI (the Julia shim throws ErrorException the same way)."""

    def __init__(self, status, message):
        super().__init__("libnepb200 status %d: %s" % (status, message))
        self.status = status


class SingularException(NepbError):
    """Zero / non-finite pivot in the device LU (reference: LinearAlgebra.SingularException)."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libnepb200.so not found at %s -- build it with `make -C %s/csrc` (or __graft_entry__.build()); "
        "the nepb200 hot path has no CPU fallback" % (LIB_PATH, _HERE)
    )

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

c_i64 = C.c_int64
c_int = C.c_int
c_dbl = C.c_double
vp = C.c_void_p
P = C.POINTER

# name -> (restype, argtypes); every symbol include/nepb200.h declares must appear here
# (tests/test_abi.py cross-checks this table against the header).
SIGNATURES = {
    "nepb_version": (C.c_char_p, []),
    "nepb_last_error": (C.c_char_p, []),
    "nepb_device_count": (c_int, [P(c_int)]),
    "nepb_set_device": (c_int, [c_int]),
    "nepb_set_stream": (c_int, [vp]),
    "nepb_synchronize": (c_int, []),
    "nepb_timer_start": (c_int, []),
    "nepb_timer_stop": (c_int, [P(C.c_float)]),
    "nepb_launch_count": (c_i64, []),
    "nepb_spmf_create": (c_int, [c_i64, c_int, P(vp), P(vp), P(vp), c_int, c_int, P(vp)]),
    "nepb_spmf_destroy": (c_int, [vp]),
    "nepb_spmf_info": (c_int, [vp, P(c_i64), P(c_int), P(c_i64), P(c_int)]),
    "nepb_spmf_pattern": (c_int, [vp, vp, vp]),
    "nepb_spmf_pattern_csr": (c_int, [vp, vp, vp, vp]),
    "nepb_spmf_mder": (c_int, [vp, vp, vp]),
    "nepb_spmf_apply": (c_int, [vp, c_int, c_int, c_int, vp, c_i64, vp, vp, c_i64]),
    "nepb_block_create": (c_int, [c_i64, c_int, P(vp)]),
    "nepb_block_destroy": (c_int, [vp]),
    "nepb_block_upload": (c_int, [vp, c_int, c_int, vp, c_i64]),
    "nepb_block_download": (c_int, [vp, c_int, c_int, vp, c_i64]),
    "nepb_block_dev_ptr": (vp, [vp]),
    "nepb_host_register": (c_int, [vp, c_i64]),
    "nepb_host_unregister": (c_int, [vp]),
    "nepb_spmf_apply_block": (c_int, [vp, c_int, vp, c_int, vp, vp]),
    "nepb_spmf_apply_bytes": (c_i64, [vp, c_int, c_int, c_int]),
    "nepb_spmf_tiles_info": (c_int, [vp, P(c_i64), P(c_i64), P(c_int)]),
    "nepb_spmf_tiles2d_info": (c_int, [vp, P(c_int), P(c_int), P(c_int), P(c_i64), P(c_i64)]),
    "nepb_lu_set_options": (c_int, [vp, c_int, c_int, c_int, vp]),
    "nepb_lu_symbolic_info": (c_int, [vp, P(c_i64), P(c_i64), P(c_int), P(c_int), P(c_int), P(c_dbl)]),
    "nepb_lu_symbolic_get": (c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "nepb_lu_analyse_pattern": (c_int, [c_i64, vp, vp, c_int, c_int, c_int, c_int, vp, vp, vp, vp, vp, vp, vp]),
    "nepb_lu_matching": (c_int, [c_i64, vp, vp, c_int, vp, vp, vp, vp]),
    "nepb_lu_create": (c_int, [vp, c_int, vp, P(vp)]),
    "nepb_lu_destroy": (c_int, [vp]),
    "nepb_lu_status": (c_int, [vp, c_int, P(c_int), P(c_int), P(c_dbl)]),
    "nepb_lu_solve": (c_int, [vp, c_int, c_int, vp, c_i64, vp, c_i64, c_int, P(c_dbl)]),
    "nepb_contour_create": (c_int, [vp, c_int, c_int, c_int, P(vp)]),
    "nepb_contour_destroy": (c_int, [vp]),
    "nepb_contour_integrate": (c_int, [vp, c_int, vp, vp, vp, c_i64, c_int, vp, vp]),
    "nepb_contour_set_probe": (c_int, [vp, vp, c_i64]),
    "nepb_contour_integrate_dev": (c_int, [vp, c_int, vp, vp, c_int]),
    "nepb_contour_get_moments": (c_int, [vp, vp]),
    "nepb_comm_unique_id": (c_int, [vp]),
    "nepb_comm_init": (c_int, [c_int, c_int, vp]),
    "nepb_comm_destroy": (c_int, []),
    "nepb_comm_info": (c_int, [P(c_int), P(c_int), P(c_int)]),
    "nepb_comm_allreduce_sum_dev": (c_int, [vp, c_i64]),
    "nepb_spmf_apply_block_ex": (c_int, [vp, c_int, vp, c_int, c_int, c_int, vp, vp, c_int]),
    "nepb_lu_solve_block": (c_int, [vp, c_int, vp, c_int, c_int, vp, c_int, vp]),
    "nepb_lu_solve_block_ex": (c_int, [vp, c_int, vp, c_int, c_int, vp, c_int, vp, c_int, P(c_dbl)]),
    "nepb_orth_dgks": (c_int, [vp, c_int, vp, c_int, c_i64, vp, P(c_dbl), P(c_int)]),
    "nepb_block_gemm": (c_int, [vp, c_int, c_int, vp, c_i64, c_int, vp, c_int, c_i64]),
    "nepb_block_copy_cols": (c_int, [vp, c_int, c_int, vp, c_int, vp, c_i64]),
    "nepb_iar_expand": (c_int, [vp, c_int, c_i64, c_int, vp, c_int, c_int]),
    "nepb_iar_pack": (c_int, [vp, c_int, c_int, c_i64, vp, c_int]),
    "nepb_block_colnorms": (c_int, [vp, c_int, c_int, c_i64, vp]),
    "nepb_wep_create": (c_int, [c_int, c_int, c_dbl, c_dbl, vp, vp, vp, P(vp)]),
    "nepb_wep_destroy": (c_int, [vp]),
    "nepb_wep_set_table": (c_int, [vp, c_int, vp]),
    "nepb_wep_info": (c_int, [vp, P(c_int), P(c_int), P(c_i64)]),
    "nepb_wep_mlincomb_block": (c_int, [vp, vp, vp, c_int, c_int, vp, vp, vp, c_int]),
    "nepb_wep_pinv": (c_int, [vp, vp, vp, vp]),
    "nepb_wep_schur_matvec_block": (c_int, [vp, vp, vp, vp, c_int, vp, c_int]),
    "nepb_wep_mlincomb_bytes": (c_i64, [vp, c_int]),
    "nepb_msws_init": (c_int, [C.c_uint64, C.c_uint64, vp]),
    "nepb_msws_fill": (c_int, [vp, c_i64, vp]),
}


def _bind_all():
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it: fail loudly
        fn.restype = res
        fn.argtypes = args


_bind_all()


def last_error() -> str:
    return lib.nepb_last_error().decode("utf-8", "replace")


def check(status: int):
    if status == NEPB_OK:
        return
    msg = last_error()
    if status == NEPB_E_SINGULAR:
        raise SingularException(status, msg)
    raise NepbError(status, msg)


def ptr(a: np.ndarray):
    return a.ctypes.data_as(vp)


def as_c128_f(a) -> np.ndarray:
    """Column-major complex128 copy/view (Julia Matrix{ComplexF64} layout)."""
    return np.asfortranarray(np.asarray(a, dtype=np.complex128))


def device_count() -> int:
    n = c_int(0)
    st = lib.nepb_device_count(C.byref(n))
    if st != NEPB_OK:
        return 0
    return n.value


def msws_fill(state: np.ndarray, count: int) -> np.ndarray:
    out = np.empty(count, dtype=np.float64)
    check(lib.nepb_msws_fill(ptr(state), count, ptr(out)))
    return out


def msws_state(seed: int = 0) -> np.ndarray:
    st = np.zeros(6, dtype=np.uint64)
    check(lib.nepb_msws_init(seed & ((1 << 64) - 1), seed >> 64, ptr(st)))
    return st
