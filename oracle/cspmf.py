"""ctypes loader of the C/OpenMP restatement of compute_MM (oracle/csrc/spmf_mm.c) -- test / baseline infrastructure only."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_spmf.so")


def _load():
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    for name in ("oracle_spmf_mm_csc", "oracle_spmf_mm_csr"):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.c_int64, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), vp, C.c_int, vp, vp, C.c_int]
    lib.oracle_max_threads.restype = C.c_int
    return lib


def _arrs(mats, fmt):
    ptrs, idxs, vals = [], [], []
    for A in mats:
        A = A.tocsc() if fmt == "csc" else A.tocsr()
        A.sort_indices()
        ptrs.append(np.ascontiguousarray(A.indptr, dtype=np.int64))
        idxs.append(np.ascontiguousarray(A.indices, dtype=np.int64))
        vals.append(np.ascontiguousarray(A.data, dtype=np.float64))
    return ptrs, idxs, vals


class CSpmf:
    """Z = sum_i f_i * A_i * V on the host cores: `csc` = the reference's loop (NEPTypes.jl:296-316 over SparseArrays' CSC
    product), dense columns over threads; `csr` = row-parallel all-core variant."""

    def __init__(self, mats):
        self.lib = _load()
        self.n = mats[0].shape[0]
        self.p = len(mats)
        self.csc = _arrs(mats, "csc")
        self.csr = None
        self._mats = mats

    def max_threads(self):
        return int(self.lib.oracle_max_threads())

    def _call(self, fn, arrs, f, V, threads):
        vp = C.c_void_p
        f = np.ascontiguousarray(np.asarray(f, dtype=np.complex128))
        V = np.asfortranarray(np.asarray(V, dtype=np.complex128))
        if V.ndim == 1:
            V = V.reshape(-1, 1, order="F")
        Z = np.empty_like(V, order="F")
        mk = lambda lst: (vp * self.p)(*[a.ctypes.data for a in lst])
        rc = fn(self.n, self.p, mk(arrs[0]), mk(arrs[1]), mk(arrs[2]), f.ctypes.data, V.shape[1], V.ctypes.data, Z.ctypes.data, threads)
        assert rc == 0
        return Z

    def mm_csc(self, f, V, threads=1):
        return self._call(self.lib.oracle_spmf_mm_csc, self.csc, f, V, threads)

    def mm_csr(self, f, V, threads=1):
        if self.csr is None:
            self.csr = _arrs(self._mats, "csr")
        return self._call(self.lib.oracle_spmf_mm_csr, self.csr, f, V, threads)
