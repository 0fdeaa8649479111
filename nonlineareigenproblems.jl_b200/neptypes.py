"""Host-side mirror of the reference's NEP plugin interface for the SPMF hot path.

Mirrors (reference file:line, relative to src/):
  NEP contract            NEPCore.jl:55-69 (size), :89 (compute_Mder), :113-160 (compute_Mlincomb[!] incl.
                          the a-vector and startder conventions), :192 (compute_MM), :272 (compute_resnorm)
  AbstractSPMF            NEPTypes.jl:96-113 (get_Av / get_fv)
  SPMF_NEP / PEP / DEP / SumNEP descriptors   NEPTypes.jl:162-237, types_poly.jl:31-41, NEPTypes.jl:427-441, :838-898

`B200SPMF` is the drop-in: it holds a `nepb_spmf` handle (union-pattern CSR with interleaved term
values, resident in HBM) and answers the whole compute contract from the fused multi-term SpMM kernel.
All f_i evaluations (scalars, bidiagonal columns, small matrix functions) stay on the host exactly as
in the reference; only coefficient blocks cross the C ABI.  Everything is ComplexF64; other element
types are the Julia shim's business (they fall through to the reference methods there).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import lib, check, ptr
from .functions import ScalarFunction, Monomial, Exp, as_function, ONE


# ---------------------------------------------------------------------------------------------
# problem descriptors (host only; they just carry matrices + functions, like the Julia structs)
# ---------------------------------------------------------------------------------------------
class SPMF_NEP:
    def __init__(self, A, fi):
        if len(A) != len(fi):
            raise ValueError("Inconsistency: Number of supplied matrices = %d but the number of supplied functions are = %d"
                             % (len(A), len(fi)))
        for a in A[1:]:
            if a.shape != A[0].shape:
                raise ValueError("The dimensions of the matrices mismatch")
        self.A = list(A)
        self.fi = [as_function(f) for f in fi]
        self.n = A[0].shape[0]

    def get_Av(self):
        return self.A

    def get_fv(self):
        return self.fi


class PEP(SPMF_NEP):
    """M(l) = sum_i A_i l^i (types_poly.jl:31-41, get_fv :83-98)."""

    def __init__(self, A):
        super().__init__(A, [Monomial(i) for i in range(len(A))])


class DEP(SPMF_NEP):
    """M(l) = -l I + sum_j A_j exp(-tau_j l) (NEPTypes.jl:427-441; get_Av/get_fv :485-513)."""

    def __init__(self, A, tauv=(0.0, 1.0)):
        tauv = np.asarray(tauv)
        if np.iscomplexobj(tauv) and np.any(tauv.imag != 0):
            raise ValueError("Incorrect construction of DEP. The delays need to be real.")
        n = A[0].shape[0]
        eye = sp.identity(n, format="csc") if sp.issparse(A[0]) else np.eye(n)
        fv = [Monomial(1, -1.0)] + [ONE if t == 0 else Exp(-float(t)) for t in tauv.real]
        super().__init__([eye] + list(A), fv)
        self.tauv = tauv.real.astype(np.float64)


class SumNEP(SPMF_NEP):
    """SPMFSumNEP: concatenated terms (NEPTypes.jl:896-906)."""

    def __init__(self, nep1, nep2):
        super().__init__(list(nep1.get_Av()) + list(nep2.get_Av()), list(nep1.get_fv()) + list(nep2.get_fv()))
        self.nep1, self.nep2 = nep1, nep2


class LowRankFactorizedNEP(SPMF_NEP):
    """SPMF whose terms carry factors A_i = L_i U_i^T, r = sum of the ranks (low_rank_nep.jl:24-43).  Built from the factors
    (`LowRankFactorizedNEP(L, U, f)`) or from the matrices, whose factors nleigs then derives (rk_nep.jl:66-95).  All compute
    functions are those of the SPMF (:45-60); only nleigs looks at the factors."""

    def __init__(self, A, fi, L=None, U=None):
        super().__init__(A, fi)
        if L is None:
            from .rk_helper import low_rank_lu_factors
            LU = [low_rank_lu_factors(a) for a in A]
            L, U = [x[0] for x in LU], [x[1] for x in LU]
        self.L, self.U = list(L), list(U)
        self.r = int(sum(u.shape[1] for u in self.U))

    @classmethod
    def from_factors(cls, L, U, fi):
        L = [sp.csc_matrix(x) for x in L]
        U = [sp.csc_matrix(x) for x in U]
        return cls([sp.csc_matrix(l @ u.conj().T) for l, u in zip(L, U)], fi, L, U)


# ---------------------------------------------------------------------------------------------
# device-resident dense block (n x k ComplexF64, row-major in HBM)
# ---------------------------------------------------------------------------------------------
class Block:
    def __init__(self, n, k):
        h = C.c_void_p()
        check(lib.nepb_block_create(n, k, C.byref(h)))
        self._h = h
        self.n, self.k = n, k

    @classmethod
    def from_host(cls, V):
        V = _lib.as_c128_f(V)
        if V.ndim == 1:
            V = V.reshape(-1, 1, order="F")
        b = cls(V.shape[0], V.shape[1])
        b.upload(V)
        return b

    def upload(self, V, k0=0):
        V = _lib.as_c128_f(V)
        if V.ndim == 1:
            V = V.reshape(-1, 1, order="F")
        check(lib.nepb_block_upload(self._h, k0, V.shape[1], ptr(V), V.shape[0]))

    def download(self, k0=0, kc=None):
        kc = self.k - k0 if kc is None else kc
        out = np.empty((self.n, kc), dtype=np.complex128, order="F")
        check(lib.nepb_block_download(self._h, k0, kc, ptr(out), self.n))
        return out

    def dev_ptr(self):
        return lib.nepb_block_dev_ptr(self._h)

    def close(self):
        if getattr(self, "_h", None):
            lib.nepb_block_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


# ---------------------------------------------------------------------------------------------
# the drop-in SPMF operator
# ---------------------------------------------------------------------------------------------
def _csc_arrays(A):
    """Julia SparseMatrixCSC layout: int64 colptr/rowval (0-based here), sorted rows; dense -> full CSC."""
    if sp.issparse(A):
        A = A.tocsc()
        if not A.has_sorted_indices:
            A = A.sorted_indices()
        return A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data
    A = np.asarray(A)
    n, m = A.shape
    colptr = np.arange(0, n * m + 1, n, dtype=np.int64)
    rowval = np.tile(np.arange(n, dtype=np.int64), m)
    return colptr, rowval, np.asfortranarray(A).reshape(-1, order="F")


class B200SPMF:
    """AbstractSPMF whose compute functions run on the B200 (fused multi-term CSR SpMM)."""

    def __init__(self, A, fi):
        if len(A) != len(fi):
            raise ValueError("Inconsistency: Number of supplied matrices = %d but the number of supplied functions are = %d"
                             % (len(A), len(fi)))
        self.A = list(A)
        self.fi = [as_function(f) for f in fi]
        self.n = A[0].shape[0]
        self.p = len(A)
        for a in A:
            if a.shape != (self.n, self.n):
                raise ValueError("The dimensions of the matrices mismatch")
        cplx = any(np.iscomplexobj(a.data if sp.issparse(a) else a) for a in A)
        vt = np.complex128 if cplx else np.float64
        keep = []
        cp, rv, nz = (C.c_void_p * self.p)(), (C.c_void_p * self.p)(), (C.c_void_p * self.p)()
        for i, a in enumerate(A):
            colptr, rowval, vals = _csc_arrays(a)
            vals = np.ascontiguousarray(vals, dtype=vt)
            keep += [colptr, rowval, vals]
            cp[i], rv[i], nz[i] = colptr.ctypes.data, rowval.ctypes.data, vals.ctypes.data
        h = C.c_void_p()
        check(lib.nepb_spmf_create(self.n, self.p, cp, rv, nz, 1 if cplx else 0, 0, C.byref(h)))
        self._h = h
        self.is_complex = cplx
        nnz = C.c_int64()
        check(lib.nepb_spmf_info(h, None, None, C.byref(nnz), None))
        self.nnz_union = nnz.value
        self._pattern = None

    @classmethod
    def from_nep(cls, nep):
        """Any AbstractSPMF-like object (get_Av / get_fv): SPMF_NEP, PEP, DEP, SumNEP."""
        dev = cls(nep.get_Av(), nep.get_fv())
        dev.source = nep  # the host descriptor (PEP / SumNEP / ...) decides nleigs' polynomial degree, rk_nep.jl:101-126
        return dev

    # -- AbstractSPMF ---------------------------------------------------------------------------
    def get_Av(self):
        return self.A

    def get_fv(self):
        return self.fi

    def size(self, d=None):
        return (self.n, self.n) if d is None else self.n

    def issparse(self):
        return sp.issparse(self.A[0])

    def close(self):
        if getattr(self, "_h", None):
            lib.nepb_spmf_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    # -- integer structure ------------------------------------------------------------------------
    def pattern(self):
        """Union pattern as (colptr, rowval), CSC, 0-based int64 (form_aligned_sparsity_patterns, NEPTypes.jl:244-274)."""
        if self._pattern is None:
            colptr = np.empty(self.n + 1, dtype=np.int64)
            rowval = np.empty(self.nnz_union, dtype=np.int64)
            check(lib.nepb_spmf_pattern(self._h, ptr(colptr), ptr(rowval)))
            self._pattern = (colptr, rowval)
        return self._pattern

    def pattern_csr(self):
        rowptr = np.empty(self.n + 1, dtype=np.int32)
        colind = np.empty(self.nnz_union, dtype=np.int32)
        perm = np.empty(self.nnz_union, dtype=np.int32)
        check(lib.nepb_spmf_pattern_csr(self._h, ptr(rowptr), ptr(colind), ptr(perm)))
        return rowptr, colind, perm

    # -- coefficients (host, scalar work) -----------------------------------------------------------
    def coefficients(self, lam, der=0):
        """c_i = f_i^{(der)}(lam) (NEPTypes.jl:322-332, :370-394)."""
        if der == 0:
            return np.array([complex(f(complex(lam))) for f in self.fi], dtype=np.complex128)
        return np.array([complex(f.derivative(lam, der)) for f in self.fi], dtype=np.complex128)

    # -- compute_Mder (NEPTypes.jl:336-367) ---------------------------------------------------------
    def compute_Mder(self, lam, i=0):
        c = self.coefficients(lam, i)
        colptr, rowval = self.pattern()
        nz = np.empty(self.nnz_union, dtype=np.complex128)
        check(lib.nepb_spmf_mder(self._h, ptr(c), ptr(nz)))
        M = sp.csc_matrix((nz, rowval.copy(), colptr.copy()), shape=(self.n, self.n))
        return M if self.issparse() else M.toarray()

    # -- raw kernel access ----------------------------------------------------------------------------
    def apply(self, mode, V, Cblk, q, out=None):
        """Z = sum_i A_i (V C_i), host arrays in / out (`out`: optional preallocated n x q column-major result)."""
        V = _lib.as_c128_f(V)
        if V.ndim == 1:
            V = V.reshape(-1, 1, order="F")
        n, k = V.shape
        if n != self.n:
            raise ValueError("V has %d rows, the NEP has size %d" % (n, self.n))
        Cblk = np.ascontiguousarray(Cblk, dtype=np.complex128)
        if out is None:
            Z = np.empty((n, q), dtype=np.complex128, order="F")
        else:
            Z = out
            if Z.shape != (n, q) or Z.dtype != np.complex128 or not Z.flags.f_contiguous:
                raise ValueError("`out` must be a column-major complex128 array of shape (%d, %d)" % (n, q))
        check(lib.nepb_spmf_apply(self._h, mode, k, q, ptr(V), n, ptr(Cblk), ptr(Z), n))
        return Z

    def apply_block(self, mode, Vb: Block, Cblk, Zb: Block):
        Cblk = np.ascontiguousarray(Cblk, dtype=np.complex128)
        check(lib.nepb_spmf_apply_block(self._h, mode, Vb._h, Zb.k, ptr(Cblk), Zb._h))

    def tiles2d_info(self):
        """(line, segments, rows per segment, tiles, staged V rows in total) of the two-dimensional tiles; line = 0: not applicable."""
        ln, sg, sr, nt, tot = C.c_int(), C.c_int(), C.c_int(), C.c_int64(), C.c_int64()
        check(lib.nepb_spmf_tiles2d_info(self._h, C.byref(ln), C.byref(sg), C.byref(sr), C.byref(nt), C.byref(tot)))
        return ln.value, sg.value, sr.value, nt.value, tot.value

    def tiles_info(self):
        """Row tiles of the multi-column kernel: (tiles, V rows staged per product, largest tile)."""
        nt, tot, mx = C.c_int64(), C.c_int64(), C.c_int()
        check(lib.nepb_spmf_tiles_info(self._h, C.byref(nt), C.byref(tot), C.byref(mx)))
        return nt.value, tot.value, mx.value

    def apply_bytes(self, mode, k, q):
        return lib.nepb_spmf_apply_bytes(self._h, mode, k, q)

    # -- compute_Mlincomb (NEPCore.jl:113-160; SPMF NEPTypes.jl:972-1011) ------------------------------
    def lincomb_coefficients(self, lam, a):
        """Coefficient block C (p x k): C[i, j] = a_1 * f_i(S)[j, 0] with S = bidiag(lam; (a_{j+1}/a_j) j),
        after the zero-entry convention of NEPTypes.jl:982-983.  Returns (C, zero_mask)."""
        a = np.array(a, dtype=np.complex128, copy=True)
        k = len(a)
        zero = a == 0
        a[zero] = 1
        s = (a[1:] / a[:-1]) * np.arange(1, k)
        Cm = np.empty((self.p, k), dtype=np.complex128)
        for i, f in enumerate(self.fi):
            Cm[i, :] = f.bidiag_column(lam, s) * a[0]
        Cm[:, zero] = 0  # zeroed columns of V == zero coefficient
        return Cm, zero

    def compute_Mlincomb(self, lam, V, a=None, startder=None):
        V = np.asarray(V)
        vec = V.ndim == 1
        Vm = V.reshape(self.n, -1, order="F") if vec else V
        k = Vm.shape[1]
        a = np.ones(k, dtype=np.complex128) if a is None else np.asarray(a, dtype=np.complex128)
        if len(a) != k:
            raise ValueError("a has %d entries for %d columns" % (len(a), k))
        if startder:  # NEPCore.jl:156-160: zero-padding == leading zero coefficients
            a = np.concatenate([np.zeros(startder, dtype=np.complex128), a])
            Cm, _ = self.lincomb_coefficients(lam, a)
            Cm = Cm[:, startder:]
        else:
            Cm, _ = self.lincomb_coefficients(lam, a)
        # GENERAL mode, q = 1: C_i is the k-vector Cm[i, :]
        z = self.apply(_lib.COEF_GENERAL, Vm, Cm.reshape(self.p, k), 1)
        return z[:, 0].copy()

    # in-place variant: same result; V/a may be clobbered by the reference, we simply do not
    compute_Mlincomb_inplace = compute_Mlincomb

    # -- compute_MM (NEPTypes.jl:276-319) ----------------------------------------------------------------
    def compute_MM(self, S, V):
        S = np.atleast_2d(np.asarray(S, dtype=np.complex128))
        V = np.asarray(V)
        if V.ndim == 1:
            V = V.reshape(-1, 1)
        q = S.shape[0]
        if V.shape[1] != q:
            raise ValueError("V has %d columns, S is %d x %d" % (V.shape[1], q, q))
        d = np.diag(S)
        if np.count_nonzero(S - np.diag(d)) == 0:  # diagonal fast path (:299-311)
            Cd = np.empty((self.p, q), dtype=np.complex128)
            for i, f in enumerate(self.fi):
                Cd[i, :] = [complex(f(complex(s))) for s in d]
            if np.all(d == d[0]):
                return self.apply(_lib.COEF_SCALAR, V, Cd[:, 0].copy(), q)
            return self.apply(_lib.COEF_DIAG, V, np.asfortranarray(Cd).reshape(-1, order="F"), q)
        blocks = np.stack([np.asfortranarray(np.asarray(f(S), dtype=np.complex128)).reshape(-1, order="F") for f in self.fi])
        return self.apply(_lib.COEF_GENERAL, V, blocks, q)

    # -- M(lam) V and batched residuals (errmeasure.jl:128-130 for all Ritz pairs at once) --------------
    def apply_M(self, lam, V):
        return self.apply(_lib.COEF_SCALAR, V, self.coefficients(lam), np.atleast_2d(np.asarray(V).T).T.shape[1])

    def residual_norms(self, lams, V):
        """||M(lam_s) v_s|| / ||v_s|| for all columns s in one multi-lambda SpMM."""
        V = np.asarray(V, dtype=np.complex128)
        lams = np.asarray(lams, dtype=np.complex128)
        q = len(lams)
        Cd = np.empty((self.p, q), dtype=np.complex128)
        for i, f in enumerate(self.fi):
            Cd[i, :] = [complex(f(complex(s))) for s in lams]
        R = self.apply(_lib.COEF_DIAG, V, np.asfortranarray(Cd).reshape(-1, order="F"), q)
        return np.linalg.norm(R, axis=0) / np.linalg.norm(V, axis=0)

    def compute_resnorm(self, lam, v):
        return float(np.linalg.norm(self.compute_Mlincomb(lam, v)))


# ---------------------------------------------------------------------------------------------
# projection W^H M(lam) V (Proj_SPMF_NEP, NEPTypes.jl:652-800; create_proj_NEP :600-640)
# ---------------------------------------------------------------------------------------------
class B200ProjSPMF:
    """Proj_SPMF_NEP for a device operator: N(lam) = sum_i f_i(lam) B_i with B_i = W^H A_i V.
    All A_i V come from ONE fused pass over the operator (GENERAL mode with selector blocks: output columns i*k..(i+1)*k-1
    are A_i V), instead of p sparse products that each re-read their own index arrays (:733-736); the k x k matrices B_i
    and the compute functions of the projected problem live on the host, as in the reference (it builds a small dense
    SPMF_NEP and delegates to it, :793-800)."""

    def __init__(self, nep: B200SPMF, maxsize=None):
        self.orgnep = nep
        self.fi = nep.get_fv()
        self.maxsize = maxsize
        self.B = [np.zeros((0, 0), dtype=np.complex128) for _ in self.fi]

    def _terms_times(self, V):
        """[A_1 V, ..., A_p V] as an n x (p k) array from one fused product."""
        V = np.asarray(V, dtype=np.complex128)
        if V.ndim == 1:
            V = V.reshape(-1, 1)
        k, p = V.shape[1], len(self.fi)
        blocks = np.zeros((p, k, p * k), dtype=np.complex128)
        for t in range(p):
            blocks[t, np.arange(k), t * k + np.arange(k)] = 1.0  # C_t selects term t into its own column window
        flat = np.stack([np.asfortranarray(blocks[t]).reshape(-1, order="F") for t in range(p)])
        return self.orgnep.apply(_lib.COEF_GENERAL, V, flat, p * k), k

    def set_projectmatrices(self, W, V):
        W = np.asarray(W, dtype=np.complex128)
        Z, k = self._terms_times(V)
        if self.maxsize is not None and k > self.maxsize:
            raise ValueError("projection larger than the preallocated size")  # the @assert of :729
        WT = W.conj().T
        self.B = [WT @ Z[:, t * k:(t + 1) * k] for t in range(len(self.fi))]
        return self

    def expand_projectmatrices(self, Wnew, Vnew):
        """Only the last row and column of every B_i are new (:774-791): two fused passes with 1 and k+1 columns."""
        Wnew, Vnew = np.asarray(Wnew, dtype=np.complex128), np.asarray(Vnew, dtype=np.complex128)
        k = Vnew.shape[1] - 1
        Zv, _ = self._terms_times(Vnew[:, -1])        # A_i v
        Zr, kk = self._terms_times(Vnew[:, :k + 1])   # A_i [V v] for the new row
        WT = Wnew[:, :k].conj().T
        w = Wnew[:, -1].conj()
        out = []
        for t, Bold in enumerate(self.B):
            Bn = np.zeros((k + 1, k + 1), dtype=np.complex128)
            Bn[:k, :k] = Bold[:k, :k]
            Bn[:k, k] = WT @ Zv[:, t]
            Bn[k, :] = w @ Zr[:, t * kk:(t + 1) * kk]
            out.append(Bn)
        self.B = out
        return self

    # the projected problem is a k x k dense SPMF evaluated on the host (:793-800)
    def compute_Mder(self, lam, der=0):
        return sum(f.derivative(lam, der) * B for f, B in zip(self.fi, self.B))

    def compute_MM(self, S, V):
        S = np.atleast_2d(np.asarray(S, dtype=np.complex128))
        return sum(B @ np.asarray(V, dtype=np.complex128) @ np.asarray(f(S), dtype=np.complex128) for f, B in zip(self.fi, self.B))

    def compute_Mlincomb(self, lam, V, a=None):
        V = np.asarray(V, dtype=np.complex128)
        Vm = V.reshape(V.shape[0], -1)
        k = Vm.shape[1]
        a = np.ones(k, dtype=np.complex128) if a is None else np.asarray(a, dtype=np.complex128)
        return sum(B @ (Vm @ (a * np.array([f.derivative(lam, j) for j in range(k)]))) for f, B in zip(self.fi, self.B))


def create_proj_NEP(nep: B200SPMF, maxsize=None):
    return B200ProjSPMF(nep, maxsize)
