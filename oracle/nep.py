"""Oracle: NEP types and the compute contract (test infrastructure, see oracle/__init__.py).

NumPy/SciPy restatement of
  * src/NEPCore.jl:113-160 (a-scaling default, startder zero padding), :212-263 (from_MM / from_Mder),
  * src/NEPTypes.jl:162-394 (SPMF_NEP: ctor/alignment, compute_MM, compute_Mder),
    :427-513 (DEP), :838-898 (SumNEP), :940-1045 (compute_Mlincomb for DEP/SPMF/PEP),
    :1055-1160 (DerSPMF),
  * src/types_poly.jl:31-98 (PEP).
Functions f_i are Python callables valid for scalars and for square matrices, as in the reference.
All arithmetic is complex128/float64 (the reference's default ComplexF64).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp


# --------------------------------------------------------------------------------------------
# scalar-or-matrix functions (the reference passes Julia closures such as S -> exp(-t*S))
# --------------------------------------------------------------------------------------------
def _is_mat(S):
    return isinstance(S, np.ndarray) and S.ndim == 2


def f_one(S):
    return np.eye(S.shape[0], dtype=S.dtype) if _is_mat(S) else 1.0


def f_id(S):
    return S


def f_neg(S):
    return -S


def f_pow(j):
    def f(S):
        return np.linalg.matrix_power(S, j) if _is_mat(S) else S ** j
    return f


def _jordan_like(S):
    """True when S = lam*I + (strictly lower bidiagonal): the only matrices compute_Mlincomb / DerSPMF feed to f_i
    (NEPTypes.jl:993, :1117-1121).  For those f(S) = sum_d t_d N^d exactly (N nilpotent), t_d the Taylor coefficients."""
    k = S.shape[0]
    if k < 2:
        return False
    d = np.diag(S)
    if np.any(d != d[0]):
        return False
    R = S - np.diag(d) - np.diag(np.diag(S, -1), -1)
    return not np.any(R)


def _fun_of_jordan_like(S, t0, ratio):
    """f(S) for S = lam*I + subdiag(s): entry (i, i-d) = t_d * prod_{l=i-d+1..i} s_l, built from the Taylor ratios
    rho_d = t_d / t_{d-1} so that no intermediate under/overflows (same quantity the reference obtains from its
    generic matrix function; checked against sqrtm/expm at small sizes in tests/test_oracle_golden.py)."""
    k = S.shape[0]
    s = np.diag(S, -1).astype(np.complex128)
    F = np.zeros((k, k), dtype=np.complex128)
    for i in range(k):
        e = t0
        F[i, i] = e
        for d in range(1, i + 1):
            e = e * ratio(d) * s[i - d]
            F[i, i - d] = e
    return F


def f_exp(c):
    """S -> exp(c*S)."""
    def f(S):
        if _is_mat(S):
            if _jordan_like(S) and S.shape[0] > 6:
                return _fun_of_jordan_like(S, np.exp(c * S[0, 0]), lambda d: c / d)
            return sla.expm(c * S)
        return np.exp(c * S)
    return f


def f_isqrt_shift(c):
    """S -> 1im*sqrt(S - c*I): the gun nonlinearities (NLEVP_native.jl:13-14), principal branch."""
    def f(S):
        if _is_mat(S):
            if _jordan_like(S) and S.shape[0] > 6:
                z = complex(S[0, 0]) - c
                return _fun_of_jordan_like(S, 1j * np.sqrt(z), lambda d: (0.5 - d + 1) / (d * z))
            return 1j * sla.sqrtm(S.astype(np.complex128) - c * np.eye(S.shape[0]))
        return 1j * np.sqrt(complex(S) - c)
    return f


# --------------------------------------------------------------------------------------------
# types
# --------------------------------------------------------------------------------------------
class SPMF_NEP:
    """NEPTypes.jl:162-237. `align_sparsity_patterns=True` follows form_aligned_sparsity_patterns (:244-274)."""

    def __init__(self, A, fi, align_sparsity_patterns=False):
        if len(A) != len(fi):
            raise ValueError("Inconsistency: Number of supplied matrices = %d but the number of supplied functions are = %d" % (len(A), len(fi)))
        sparse_flags = [sp.issparse(a) for a in A]
        if not (all(sparse_flags) or not any(sparse_flags)):
            raise ValueError("Mixing sparse and dense matrices is not allowed in SPMF_NEP.")
        for a in A[1:]:
            if a.shape != A[0].shape:
                raise ValueError("The dimensions of the matrices mismatch")
        self.n = A[0].shape[0]
        self.fi = list(fi)
        self.sparsity_patterns_aligned = False
        if all(sparse_flags) and align_sparsity_patterns:
            A = form_aligned_sparsity_patterns(A)
            self.sparsity_patterns_aligned = True
        self.A = list(A)


def form_aligned_sparsity_patterns(AA):
    """NEPTypes.jl:244-274: every matrix re-expressed on the union pattern (CSC, sorted rows)."""
    Zero = None
    for A in AA:
        P = sp.csc_matrix((np.ones(A.nnz), A.tocsc().indices, A.tocsc().indptr), shape=A.shape)
        Zero = P if Zero is None else Zero + P
    Zero = Zero.tocsc()
    Zero.sort_indices()
    out = []
    for A in AA:
        A = A.tocsc()
        S = sp.csc_matrix((np.zeros(Zero.nnz, dtype=A.dtype), Zero.indices.copy(), Zero.indptr.copy()), shape=A.shape)
        for col in range(A.shape[1]):
            lo, hi = Zero.indptr[col], Zero.indptr[col + 1]
            rows_u = Zero.indices[lo:hi]
            a_lo, a_hi = A.indptr[col], A.indptr[col + 1]
            pos = np.searchsorted(rows_u, A.indices[a_lo:a_hi])
            np.add.at(S.data, lo + pos, A.data[a_lo:a_hi])
        out.append(S)
    return out


class PEP:
    """types_poly.jl:31-41: M(l) = sum_i A_i l^i."""

    def __init__(self, A):
        self.A = list(A)
        self.n = A[0].shape[0]


class DEP:
    """NEPTypes.jl:427-441: M(l) = -l I + sum_j A_j exp(-tau_j l)."""

    def __init__(self, A, tauv=(0.0, 1.0)):
        self.A = list(A)
        self.n = A[0].shape[0]
        tauv = np.asarray(tauv)
        if np.iscomplexobj(tauv) and np.any(tauv.imag != 0):
            raise ValueError("Incorrect construction of DEP. The delays need to be real.")
        self.tauv = np.asarray(tauv.real, dtype=np.float64)


class SumNEP:
    """NEPTypes.jl:838-898 (SPMFSumNEP)."""

    def __init__(self, nep1, nep2):
        self.nep1, self.nep2 = nep1, nep2
        self.n = nep1.n


class DerSPMF:
    """NEPTypes.jl:1055-1128: table fD[j,i] = f_i^{(j)}(sigma), j = 0..2m+1, through f_i(SS)[:,0] with
    SS = diag(sigma) + subdiag(1..2m+1)."""

    def __init__(self, spmf, sigma, m):
        self.spmf = spmf
        self.sigma = sigma
        self.n = spmf.n
        fv = get_fv(spmf)
        SS = np.diag(np.full(2 * m + 2, sigma, dtype=np.complex128)) + np.diag(np.arange(1, 2 * m + 2, dtype=np.complex128), -1)
        self.fD = np.stack([np.asarray(f(SS))[:, 0] for f in fv], axis=1)


class Proj_SPMF_NEP:
    """NEPTypes.jl:652-800: N(lam) = W^H M(lam) V for an AbstractSPMF, kept as the small dense SPMF sum_i f_i(lam) (W^H A_i V).
    The compute functions of the projected problem are those of `nep_proj` (delegation, :793-800)."""

    def __init__(self, nep, maxsize=None):
        self.orgnep = nep
        self.orgnep_Av = get_Av(nep)
        self.orgnep_fv = get_fv(nep)
        self.maxsize = maxsize
        self.B = [np.zeros((0, 0), dtype=np.complex128) for _ in self.orgnep_Av]
        self.nep_proj = None

    def set_projectmatrices(self, W, V):
        """:723-740."""
        W, V = np.asarray(W), np.asarray(V)
        if self.maxsize is not None:
            assert V.shape[1] <= self.maxsize
        WT = W.conj().T
        self.B = [np.asarray(WT @ _dot(A, V), dtype=np.complex128) for A in self.orgnep_Av]
        self.nep_proj = SPMF_NEP(self.B, self.orgnep_fv)

    def expand_projectmatrices(self, Wnew, Vnew):
        """:774-791: only the new last row and column of every W^H A_i V are computed."""
        Wnew, Vnew = np.asarray(Wnew), np.asarray(Vnew)
        k = Vnew.shape[1] - 1
        w, v = Wnew[:, -1], Vnew[:, -1]
        WT = Wnew[:, :k].conj().T
        out = []
        for A, Bold in zip(self.orgnep_Av, self.B):
            Bn = np.zeros((k + 1, k + 1), dtype=np.complex128)
            Bn[:k, :k] = Bold[:k, :k]
            Bn[:k, k] = WT @ _dot(A, v)
            Bn[k, :] = w.conj() @ _dot(A, Vnew[:, :k + 1])
            out.append(Bn)
        self.B = out
        self.nep_proj = SPMF_NEP(self.B, self.orgnep_fv)


def create_proj_NEP(nep, maxsize=None):
    """NEPTypes.jl:600-640 for the SPMF case."""
    return Proj_SPMF_NEP(nep, maxsize)


def size(nep, d=None):
    return (nep.n, nep.n) if d is None else nep.n


def issparse(nep):
    return sp.issparse(get_Av(nep)[0])


def get_Av(nep):
    """NEPTypes.jl:896-906; DEP :485-493 (identity first); PEP types_poly.jl:79-81."""
    if isinstance(nep, DEP):
        n = nep.n
        J = sp.identity(n, format="csc") if sp.issparse(nep.A[0]) else np.eye(n)
        return [J] + list(nep.A)
    if isinstance(nep, SumNEP):
        return get_Av(nep.nep1) + get_Av(nep.nep2)
    if isinstance(nep, DerSPMF):
        return get_Av(nep.spmf)
    return nep.A


def get_fv(nep):
    """DEP NEPTypes.jl:495-513; PEP types_poly.jl:83-98."""
    if isinstance(nep, DEP):
        fv = [f_neg]
        for tau in nep.tauv:
            fv.append(f_one if tau == 0 else f_exp(-tau))
        return fv
    if isinstance(nep, PEP):
        fv = []
        for i in range(len(nep.A)):
            fv.append(f_one if i == 0 else (f_id if i == 1 else f_pow(i)))
        return fv
    if isinstance(nep, SumNEP):
        return get_fv(nep.nep1) + get_fv(nep.nep2)
    if isinstance(nep, DerSPMF):
        return get_fv(nep.spmf)
    return nep.fi


def _dot(A, X):
    return A @ X


# --------------------------------------------------------------------------------------------
# compute_Mder
# --------------------------------------------------------------------------------------------
def compute_Mder(nep, lam, i=0):
    if isinstance(nep, SumNEP):  # NEPTypes.jl:891-892
        return compute_Mder(nep.nep1, lam, i) + compute_Mder(nep.nep2, lam, i)
    if isinstance(nep, DerSPMF):
        return compute_Mder(nep.spmf, lam, i)
    if isinstance(nep, PEP):  # types_poly.jl:65-76
        Z = None
        for j in range(i, len(nep.A)):
            c = lam ** (j - i) * (math.factorial(j) / math.factorial(j - i))
            T = nep.A[j] * c
            Z = T if Z is None else Z + T
        if Z is None:
            Z = nep.A[0] * 0
        return Z
    if isinstance(nep, DEP):  # NEPTypes.jl:446-467
        n = nep.n
        J = sp.identity(n, format="csc", dtype=np.complex128) if sp.issparse(nep.A[0]) else np.eye(n, dtype=np.complex128)
        M = J * 0
        if i == 0:
            M = -lam * J
        if i == 1:
            M = -1.0 * J
        for j, tau in enumerate(nep.tauv):
            a = np.exp(-tau * lam) * (-tau) ** i
            M = M + nep.A[j] * a
        return M
    # SPMF_NEP, NEPTypes.jl:322-394
    if i == 0:
        x = [f(lam) for f in nep.fi]
    else:
        k = i + 1  # Jordan matrix trick (:376-386)
        S = np.diag(np.full(k, lam, dtype=np.complex128)) + np.diag(np.arange(1, k, dtype=np.complex128), -1)
        x = [np.asarray(f(S))[-1, 0] for f in nep.fi]
    Z = None
    for A, c in zip(nep.A, x):
        T = A * c
        Z = T if Z is None else Z + T
    return Z


# --------------------------------------------------------------------------------------------
# compute_MM
# --------------------------------------------------------------------------------------------
def compute_MM(nep, S, V):
    S = np.atleast_2d(np.asarray(S))
    V = np.asarray(V)
    if isinstance(nep, SumNEP):  # NEPTypes.jl:893-894
        return compute_MM(nep.nep1, S, V) + compute_MM(nep.nep2, S, V)
    if isinstance(nep, DerSPMF):
        return compute_MM(nep.spmf, S, V)
    if isinstance(nep, PEP):  # types_poly.jl:44-59
        Z = np.zeros(V.shape, dtype=np.complex128)
        Si = np.eye(S.shape[0], dtype=np.complex128)
        for A in nep.A:
            Z = Z + _dot(A, V @ Si)
            Si = Si @ S
        return Z
    if isinstance(nep, DEP):  # NEPTypes.jl:473-483
        Z = -(V @ S).astype(np.complex128)
        for A, tau in zip(nep.A, nep.tauv):
            Z = Z + _dot(A, V @ sla.expm(-tau * S))
        return Z
    # SPMF_NEP, NEPTypes.jl:276-319 (with the diagonal-S fast paths)
    n, p = nep.n, S.shape[0]
    Z = np.zeros((n, p), dtype=np.complex128)
    isdiag = np.count_nonzero(S - np.diag(np.diag(S))) == 0
    for A, f in zip(nep.A, nep.fi):
        if isdiag:
            Fi = np.diag(np.array([f(s) for s in np.diag(S)], dtype=np.complex128))
        else:
            Fi = np.asarray(f(S), dtype=np.complex128)
        Z += _dot(A, V @ Fi)
    return Z


# --------------------------------------------------------------------------------------------
# compute_Mlincomb
# --------------------------------------------------------------------------------------------
def compute_Mlincomb(nep, lam, V, a=None, startder=None):
    """sum_j a_j M^{(j-1)}(lam) V[:,j] (NEPCore.jl:113-160); never modifies V or a."""
    V = np.array(V, copy=True)
    vec_input = V.ndim == 1
    k = 1 if vec_input else V.shape[1]
    if startder is not None:  # NEPCore.jl:156-160
        if a is None:
            a = np.ones(k)
        a = np.concatenate([np.zeros(startder, dtype=np.asarray(a).dtype), np.asarray(a)])
        Vm = V.reshape(nep.n, -1)
        V = np.concatenate([np.zeros((nep.n, startder), dtype=Vm.dtype), Vm], axis=1)
        return compute_Mlincomb(nep, lam, V, a)
    a = None if a is None else np.array(a, copy=True)
    if hasattr(nep, "native_Mlincomb"):  # a NEP type with its own method, e.g. WEP_FD (Waveguide.jl:324-379, oracle/wep.py)
        return nep.native_Mlincomb(lam, V, a)
    if isinstance(nep, SumNEP):
        # generic a handling (NEPCore.jl:113-125) then delegation (NEPTypes.jl:889-890)
        if a is not None and not np.all(a == 1):
            V = V * a[0] if vec_input else V * a[None, :]
        return compute_Mlincomb(nep.nep1, lam, V) + compute_Mlincomb(nep.nep2, lam, V)
    if isinstance(nep, DerSPMF):  # NEPTypes.jl:1130-1160
        if lam != nep.sigma:
            return compute_Mlincomb(nep.spmf, lam, V, a)
        if a is None:
            a = np.ones(k)
        Vm = V.reshape(nep.n, k)
        VafD = Vm @ (a[:, None] * nep.fD[:k, :])
        z = np.zeros(nep.n, dtype=np.complex128)
        for j, A in enumerate(get_Av(nep)):
            z += _dot(A, VafD[:, j])
        return z
    if a is None:
        a = np.ones(k, dtype=np.complex128)
    if isinstance(nep, DEP):  # NEPTypes.jl:940-968
        Vm = V.reshape(nep.n, k).astype(np.complex128)
        z = np.zeros(nep.n, dtype=np.complex128)
        for A, tau in zip(nep.A, nep.tauv):
            w = np.exp(-lam * tau) * (-tau) ** np.arange(k, dtype=np.float64)
            z += _dot(A, Vm @ (a * w))
        if k == 1:
            z -= a[0] * lam * Vm[:, 0]
        else:
            z += -lam * a[0] * Vm[:, 0] - a[1] * Vm[:, 1]
        return z
    if isinstance(nep, PEP):  # NEPTypes.jl:1016-1045
        Vm = V.reshape(nep.n, k).astype(np.complex128)
        z = np.zeros(nep.n, dtype=np.complex128)
        d = len(nep.A) - 1
        kk = min(k, d + 1)
        if lam == 0:
            for j in range(kk):
                z += a[j] * math.factorial(j) * _dot(nep.A[j], Vm[:, j])
        else:
            for j in range(kk):
                for i in range(j, d + 1):
                    z += a[j] * lam ** (i - j) * (math.factorial(i) / math.factorial(i - j)) * _dot(nep.A[i], Vm[:, j])
        return z
    # SPMF_NEP, NEPTypes.jl:972-1011
    Vm = V.reshape(nep.n, k).astype(np.complex128)
    a = a.astype(np.complex128)
    zero = a == 0
    Vm[:, zero] = 0
    a[zero] = 1
    z = np.zeros(nep.n, dtype=np.complex128)
    if vec_input:
        for A, f in zip(nep.A, nep.fi):
            z += _dot(A, Vm[:, 0] * f(lam))
    else:
        S = np.diag(np.full(k, lam, dtype=np.complex128))
        if k > 1:
            S = S + np.diag((a[1:] / a[:-1]) * np.arange(1, k), -1)
        for A, f in zip(nep.A, nep.fi):
            Fi1 = np.asarray(f(S), dtype=np.complex128)[:, 0]
            z += _dot(A, Vm @ Fi1)
    return a[0] * z


def compute_Mlincomb_from_MM(nep, lam, V, a):
    """NEPCore.jl:212-228."""
    V = np.array(V, dtype=np.complex128, copy=True)
    a = np.array(a, dtype=np.complex128, copy=True)
    k = V.shape[1]
    zero = a == 0
    V[:, zero] = 0
    a[zero] = 1
    S = np.diag(np.full(k, lam, dtype=np.complex128)) + np.diag((a[1:] / a[:-1]) * np.arange(1, k), -1)
    return a[0] * compute_MM(nep, S, V)[:, 0]


def compute_Mlincomb_from_Mder(nep, lam, V, a):
    """NEPCore.jl:239-248."""
    z = np.zeros(nep.n, dtype=np.complex128)
    for i in range(len(a)):
        if a[i] != 0:
            z = z + compute_Mder(nep, lam, i) @ (V[:, i] * a[i])
    return z


def compute_resnorm(nep, lam, v):
    """NEPCore.jl:272-274."""
    return np.linalg.norm(compute_Mlincomb(nep, lam, v))


# --------------------------------------------------------------------------------------------
# error measures (errmeasure.jl:128-130, :174-191)
# --------------------------------------------------------------------------------------------
def residual_errmeasure(nep):
    def est(lam, v):
        return np.linalg.norm(compute_Mlincomb(nep, lam, v)) / np.linalg.norm(v)
    return est


def standard_spmf_errmeasure(nep):
    Av, fv = get_Av(nep), get_fv(nep)
    coeffs = [np.linalg.norm(A.data) if sp.issparse(A) else np.linalg.norm(A) for A in Av]  # Frobenius

    def est(lam, v):
        denom = sum(c * abs(f(lam)) for c, f in zip(coeffs, fv))
        return np.linalg.norm(compute_Mlincomb(nep, lam, v)) / (np.linalg.norm(v) * denom)
    return est


# --------------------------------------------------------------------------------------------
# gallery NEPs
# --------------------------------------------------------------------------------------------
def nep_gallery(name, *params):
    from . import gallery as g
    if name == "dep0":
        A0, A1, tauv = g.dep0_matrices(*params)
        return DEP([A0, A1], tauv)
    if name == "pep0":  # basic_random_examples.jl:36-44
        n = params[0] if params else 200
        rng = g.MSWS_RNG()
        return PEP([g.gen_rng_mat(rng, n, n) for _ in range(3)])
    if name == "dep0_tridiag":
        A0, A1, tauv = g.dep0_tridiag_matrices(*params)
        return DEP([A0, A1], tauv)
    if name == "neuron0":
        A, tauv = g.neuron0_matrices()
        return DEP(A, tauv)
    if name == "nlevp_native_gun":  # NLEVP_native.jl:4-18
        K, M, W1, W2 = g.load_gun_matrices()
        pep = PEP([K, -M])
        sqrtnep = SPMF_NEP([W1, W2], [f_isqrt_shift(0.0), f_isqrt_shift(108.8774 ** 2)])
        return SumNEP(pep, sqrtnep)
    if name == "qdep0":  # gallery_examples.jl:75-88
        A0, A1 = g.load_qdep0_matrices()
        n = A0.shape[0]
        return SPMF_NEP([-sp.identity(n, format="csc"), A0, A1], [f_pow(2), f_one, f_exp(-1.0)])
    raise ValueError("%s not supported" % name)


# --------------------------------------------------------------------------------------------
# deflation (src/nep_deflation.jl), MM formulation: DeflatedNEPMM (:17-21), compute_MM (:183-197), compute_Mlincomb through
# compute_Mlincomb_from_MM (:199-201), compute_Mder column by column through compute_MM (compute_Mder_from_MM, NEPCore.jl).
# The product mirrors the *Generic* formulation (binomial expansion); the reference's own test "Deflation modes" (test/deflation.jl:
# 46-92) asserts that the two formulations agree -- that assertion is what the parity tests re-use.
# --------------------------------------------------------------------------------------------
class DeflatedNEPMM:
    def __init__(self, orgnep, S0, V0):
        self.orgnep = orgnep
        self.S0 = np.atleast_2d(np.asarray(S0, dtype=np.complex128))
        self.V0 = np.asarray(V0, dtype=np.complex128).reshape(orgnep.n, -1)
        self.n = orgnep.n + self.V0.shape[1]


def deflated_compute_MM(dnep, S, V):
    """nep_deflation.jl:183-197."""
    S = np.atleast_2d(np.asarray(S, dtype=np.complex128))
    V = np.asarray(V, dtype=np.complex128)
    n0, p0, p = dnep.orgnep.n, dnep.S0.shape[0], S.shape[0]
    V1, V2 = V[:n0, :], V[n0:, :]
    Stilde = np.block([[dnep.S0, V2], [np.zeros((p, p0), dtype=np.complex128), S]])
    Vtilde = np.concatenate([dnep.V0, V1], axis=1)
    R = compute_MM(dnep.orgnep, Stilde, Vtilde)
    return np.concatenate([R[:n0, p0:], dnep.V0.conj().T @ V1], axis=0)


def deflated_compute_Mlincomb(dnep, lam, V, a=None):
    """compute_Mlincomb_from_MM (NEPCore.jl:212-228) on the deflated problem."""
    V = np.array(V, dtype=np.complex128, copy=True)
    V = V.reshape(dnep.n, -1, order="F") if V.ndim == 1 else V
    k = V.shape[1]
    a = np.ones(k, dtype=np.complex128) if a is None else np.array(a, dtype=np.complex128, copy=True)
    zero = a == 0
    V[:, zero] = 0
    a[zero] = 1
    S = np.diag(np.full(k, lam, dtype=np.complex128)) + np.diag((a[1:] / a[:-1]) * np.arange(1, k), -1)
    return a[0] * deflated_compute_MM(dnep, S, V)[:, 0]


def deflated_compute_Mder(dnep, lam, der=0):
    """compute_Mder_from_MM: column j of M^(der)(lam) = der-th derivative applied to e_j, through a Jordan block of size der + 1."""
    n = dnep.n
    M = np.zeros((n, n), dtype=np.complex128)
    S = np.diag(np.full(der + 1, lam, dtype=np.complex128)) + np.diag(np.ones(der), 1)
    fact = float(np.prod(np.arange(1, der + 1))) if der else 1.0
    for j in range(n):
        V = np.zeros((n, der + 1), dtype=np.complex128)
        V[j, 0] = 1.0
        M[:, j] = deflated_compute_MM(dnep, S, V)[:, der] * fact
    return M


def normalize_schur_pair(S, V):
    """nep_deflation.jl:278-287."""
    QQ, RR = np.linalg.qr(V)
    return RR @ S @ np.linalg.inv(RR), QQ


def deflate_eigpair(nep, lam, v):
    """nep_deflation.jl:369-398 (mode :MM)."""
    v = np.asarray(v, dtype=np.complex128)
    if isinstance(nep, DeflatedNEPMM):
        n, p0 = nep.orgnep.n, nep.V0.shape[1]
        V1 = np.zeros((n, p0 + 1), dtype=np.complex128)
        S1 = np.zeros((p0 + 1, p0 + 1), dtype=np.complex128)
        V1[:, :p0] = nep.V0
        V1[:, p0] = v[:n]
        S1[:p0, :p0] = nep.S0
        S1[:, p0] = np.concatenate([v[n:], [lam]])
        S1, V1 = normalize_schur_pair(S1, V1)
        return DeflatedNEPMM(nep.orgnep, S1, V1)
    S0, V0 = normalize_schur_pair(np.array([[complex(lam)]]), v.reshape(-1, 1))
    return DeflatedNEPMM(nep, S0, V0)


# --------------------------------------------------------------------------------------------
# LowRankFactorizedNEP (src/low_rank_nep.jl:24-43) and the low-rank LU factors nleigs works with
# (src/rk_helper/rk_nep.jl:43-98: LowRankMatrixAndFunction, low_rank_lu_factors, compactlu)
# --------------------------------------------------------------------------------------------
def low_rank_lu_factors(A):
    """rk_nep.jl:66-89: LU of the bounding box of the nonzeros, trivial columns dropped; A = L U^T (the reference ignores the
    row permutation of `lu`; as there, the factors are only valid when no row exchange happens -- asserted)."""
    import scipy.linalg as L_
    A = sp.coo_matrix(A)
    n = A.shape[0]
    r0, r1, c0, c1 = A.row.min(), A.row.max(), A.col.min(), A.col.max()
    B = A.tocsr()[r0:r1 + 1, c0:c1 + 1].toarray()
    Pm, Lf, Uf = L_.lu(B)
    assert np.allclose(Pm, np.eye(len(B))), "low_rank_lu_factors: row exchanges are not supported (rk_nep.jl:80-88)"
    m = len(B)
    sel = np.array([(np.count_nonzero(Lf[i:, i]) > 1) or (np.count_nonzero(Uf[i, i:]) > 0) for i in range(m)])  # compactlu (:91-95)
    Lca = sp.lil_matrix((n, int(sel.sum())))
    Lca[r0:r1 + 1, :] = Lf[:, sel]
    Uca = sp.lil_matrix((int(sel.sum()), A.shape[1]))
    Uca[:, c0:c1 + 1] = Uf[sel, :]
    return sp.csc_matrix(Lca), sp.csc_matrix(Uca.T)


class LowRankFactorizedNEP(SPMF_NEP):
    """An SPMF whose terms come with factors A_i = L_i U_i^T; r = sum of the ranks (low_rank_nep.jl:24-29).  All compute
    functions are those of the underlying SPMF (:45-60)."""

    def __init__(self, A, fi, L=None, U=None):
        super().__init__(A, fi)
        if L is None:
            LU = [low_rank_lu_factors(a) for a in A]
            L, U = [x[0] for x in LU], [x[1] for x in LU]
        self.L, self.U = list(L), list(U)
        self.r = int(sum(u.shape[1] for u in self.U))
