"""Solver loops that call the hot path through the plugin interface (host-side mirror of the callers).

These follow the reference's drivers line by line on the host and route every heavy operation to the device:
  contour_beyn   src/method_beyncontour.jl:49-185 + src/method_contour_common.jl:61-94 -- the N quadrature nodes go to
                 nepb_contour_integrate in one call (batched factor + solve + accumulate, sharded over ranks)
  iar            src/method_iar.jl:47-184     tiar   src/method_tiar.jl:53-257
  resinv         src/method_newton.jl:142-226 (+ compute_rf, src/compute_rf_wrapper.jl:25-54)
Small dense post-processing (eig of H, SVD of the k x k moments, sorting permutations) stays on the host as in the
reference.  Exceptions mirror NEPCore.jl:324-350.
"""
from __future__ import annotations

import ctypes as C

import warnings

import numpy as np

from . import _lib
from ._lib import lib, check, ptr
from .neptypes import B200SPMF
from .linsolve import B200LinSolverCreator, B200BackslashLinSolverCreator


class NoConvergenceException(Exception):
    def __init__(self, lam, v, errmeasure, msg):
        super().__init__(msg)
        self.lam, self.v, self.errmeasure = lam, v, errmeasure


class LostOrthogonalityException(Exception):
    pass


# ---------------------------------------------------------------------------------------------
# error measures (errmeasure.jl:91-101,128-130,174-191)
# ---------------------------------------------------------------------------------------------
class ResidualErrmeasure:
    def __init__(self, nep):
        self.nep = nep

    def estimate_error(self, lam, v):
        return float(np.linalg.norm(self.nep.compute_Mlincomb(lam, v)) / np.linalg.norm(v))

    def estimate_errors(self, lams, V):
        return self.nep.residual_norms(lams, V)


class StandardSPMFErrmeasure:
    def __init__(self, nep):
        import scipy.sparse as sp
        self.nep = nep
        self.coeffs = np.array([np.linalg.norm(A.data) if sp.issparse(A) else np.linalg.norm(A) for A in nep.get_Av()])

    def _denom(self, lam):
        return float(sum(c * abs(f(complex(lam))) for c, f in zip(self.coeffs, self.nep.get_fv())))

    def estimate_error(self, lam, v):
        return float(np.linalg.norm(self.nep.compute_Mlincomb(lam, v)) / (np.linalg.norm(v) * self._denom(lam)))

    def estimate_errors(self, lams, V):
        """All Ritz pairs in one multi-lambda SpMM instead of k SpMVs (method_iar.jl:134-135)."""
        r = self.nep.residual_norms(lams, V)
        return r / np.array([self._denom(l) for l in lams])


def DefaultErrmeasure(nep):
    """errmeasure.jl:91-101: StandardSPMFErrmeasure for an AbstractSPMF, the plain residual otherwise (e.g. WEP_FD)."""
    return StandardSPMFErrmeasure(nep) if hasattr(nep, "get_Av") else ResidualErrmeasure(nep)


def _errs(errmeasure, lams, Q):
    if hasattr(errmeasure, "estimate_errors"):
        return np.asarray(errmeasure.estimate_errors(lams, Q), dtype=float)
    f = errmeasure.estimate_error if hasattr(errmeasure, "estimate_error") else errmeasure
    return np.array([f(lams[s], Q[:, s]) for s in range(len(lams))], dtype=float)


# ---------------------------------------------------------------------------------------------
# contour integration
# ---------------------------------------------------------------------------------------------
def _check_node_flags(flags):
    """Quadrature nodes whose factorisation failed make the moments wrong: the reference gets an accurate UMFPACK solve or
    a SingularException at such a node (LinSolvers.jl:116), never a silent perturbation.  bit1 = non-finite pivot, bit0 = a
    zero pivot replaced, bit2 (4) = tiny pivots lifted -- after the library's static-pivoting fallback (bit3) has been tried."""
    near = np.flatnonzero(flags & 16)
    if near.size:
        warnings.warn("%d quadrature node(s) have pivots below 1e-8 max|M_ij|: an eigenvalue lies very close to the contour" % near.size)
    bad = np.flatnonzero(flags & (1 | 2 | 4))
    if bad.size:
        raise _lib.SingularException(_lib.NEPB_E_SINGULAR,
                                     "factorisation failed at %d quadrature node(s) (first: %d, flags %d): an eigenvalue lies on the "
                                     "contour or M is singular to working precision there" % (bad.size, bad[0], int(flags[bad[0]])))


class ContourIntegrator:
    """The MatrixIntegrator seam (method_contour_common.jl:18-45) for B200 operators: all nodes in one device call."""

    def __init__(self, nep: B200SPMF, k, mg=2, batch=32):
        h = C.c_void_p()
        check(lib.nepb_contour_create(nep._h, k, mg, batch, C.byref(h)))
        self._h = h
        self.nep, self.k, self.mg, self.batch = nep, k, mg, batch

    def integrate(self, lams, weights, Vh, reduce=False):
        """S[:,:,j] = sum_i weights[i,j] M(lams[i])^-1 Vh; returns (S (n,k,mg), node_flags)."""
        lams = np.asarray(lams, dtype=np.complex128)
        nn = len(lams)
        coef = np.ascontiguousarray(np.stack([self.nep.coefficients(l) for l in lams])) if nn else np.zeros((0, self.nep.p), complex)
        W = np.ascontiguousarray(np.asarray(weights, dtype=np.complex128).reshape(nn, self.mg))
        Vh = _lib.as_c128_f(Vh)
        S = np.empty((self.nep.n, self.k, self.mg), dtype=np.complex128, order="F")
        flags = np.zeros(max(nn, 1), dtype=np.int32)
        check(lib.nepb_contour_integrate(self._h, nn, ptr(coef), ptr(W), ptr(Vh), Vh.shape[0], 1 if reduce else 0, ptr(S), ptr(flags)))
        return S, flags[:nn]

    def close(self):
        if getattr(self, "_h", None):
            lib.nepb_contour_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def beyn_extract(A0, A1, sigma, radius, k, neigs, tol, rank_drop_tol, errmeasure, sanity_check):
    """method_beyncontour.jl:113-184 (host, k x k sized work + the sorting / filtering permutations)."""
    V, S, Wh = np.linalg.svd(A0, full_matrices=False)
    p = int(np.count_nonzero(S / S[0] > rank_drop_tol))
    V0, W0 = V[:, :p], Wh.conj().T[:, :p]
    B = (V0.conj().T @ A1 @ W0) @ np.diag(1.0 / S[:p])
    lam, VB = np.linalg.eig(B)
    lam = lam + sigma
    Vv = V0 @ VB
    r1, r2 = radius
    info = {"p": p, "S": S}

    def inside_of(ls):
        return ((ls - sigma).real / r1) ** 2 + ((ls - sigma).imag / r2) ** 2 <= 1

    if not sanity_check:
        si = np.argsort(np.abs(sigma - lam), kind="stable")
        idx = si[np.argsort(~inside_of(lam[si]), kind="stable")]
        return lam[idx], Vv[:, idx], info
    errs = _errs(errmeasure, lam, Vv)
    info["errs"] = errs
    good = np.nonzero(errs < tol)[0]
    sg = good[np.argsort(np.abs(sigma - lam[good]), kind="stable")]
    idx = sg[np.argsort(~inside_of(lam[sg]), kind="stable")]
    if len(idx) > neigs:
        idx = idx[:neigs]
    return lam[idx], Vv[:, idx], info


def beyn_quadrature(N, radius, sigma, rank=0, world=1):
    """Nodes and weights of the trapezoid rule on the ellipse (method_beyncontour.jl:69-70, method_contour_common.jl:62-93)
    owned by `rank` of `world` (round-robin i = rank mod world).  Returns (indices, lambda_i = g(t_i) + sigma, W) with
    W[i, :] = (gp_i, gp_i g_i) * h / (2 pi i), so that A_j = sum_i W[i, j] M(lambda_i)^-1 Vh summed over all ranks."""
    radius = (radius, radius) if np.isscalar(radius) else tuple(radius)
    h = 2 * np.pi / N
    t = h * np.arange(N)
    g = radius[0] * np.cos(t) + 1j * radius[1] * np.sin(t)
    gp = -radius[0] * np.sin(t) + 1j * radius[1] * np.cos(t)
    W = np.stack([gp * h / (2j * np.pi), gp * g * h / (2j * np.pi)], axis=1)
    mine = np.arange(rank, N, world)
    return mine, g[mine] + sigma, W[mine]


def contour_beyn(nep: B200SPMF, Vh=None, sigma=0.0, radius=1.0, N=1000, neigs=2, k=None, tol=np.sqrt(np.finfo(float).eps),
                 errmeasure=None, sanity_check=True, rank_drop_tol=None, batch=32, rank=0, world=1, integrator=None,
                 return_moments=False, seed=10):
    """contour_beyn for B200 operators.  rank/world shard the quadrature nodes (i = rank mod world); with world > 1 the
    library communicator (nepb_comm_init) must exist and the moments are summed with one NCCL all-reduce."""
    n = nep.n
    k = neigs + 1 if k is None else k
    if k > n:
        raise ValueError("Cannot compute more eigenvalues than the size of the NEP with contour_beyn() k=%d n=%d" % (k, n))
    if k <= 0:
        raise ValueError("k must be positive, k=%d." % k)
    radius = (radius, radius) if np.isscalar(radius) else tuple(radius)
    rank_drop_tol = tol if rank_drop_tol is None else rank_drop_tol
    if Vh is None:  # the reference draws randn(n,k) after Random.seed!(10); any fixed Gaussian probe is equivalent
        Vh = np.random.default_rng(seed).standard_normal((n, k))
    errmeasure = errmeasure or DefaultErrmeasure(nep)
    mine, lams, Wm = beyn_quadrature(N, radius, sigma, rank, world)
    own = integrator is None
    integ = integrator or ContourIntegrator(nep, k, 2, min(batch, max(1, len(mine))))
    try:
        S, flags = integ.integrate(lams, Wm, Vh, reduce=world > 1)
    finally:
        if own:
            integ.close()
    _check_node_flags(flags)
    A0, A1 = S[:, :, 0], S[:, :, 1]
    lam, V, info = beyn_extract(A0, A1, sigma, radius, k, neigs, tol, rank_drop_tol, errmeasure, sanity_check)
    info["node_flags"] = flags
    if return_moments:
        return lam, V, A0, A1, info
    return lam, V


# ---------------------------------------------------------------------------------------------
# orthogonalisation (host reference implementation of the DGKS hook; the device version is in orth.py)
# ---------------------------------------------------------------------------------------------
def dgks_host(V, w, h):
    h[:] = V.conj().T @ w
    w -= V @ h
    nrm = np.linalg.norm(w)
    ps = np.linalg.norm(h)
    while nrm < ps / np.sqrt(2.0):
        c = V.conj().T @ w
        ps = np.linalg.norm(c)
        w -= V @ c
        h += c
        nrm = np.linalg.norm(w)
    w *= 1.0 / nrm
    return nrm


# ---------------------------------------------------------------------------------------------
# resinv
# ---------------------------------------------------------------------------------------------
def compute_rf(nep, x, y=None, lam=0.0, tol=np.finfo(float).eps * 100, maxit=80):
    y = x if y is None else y
    li = complex(lam)
    dl, count = np.inf, 0
    while abs(dl) > tol and count < maxit:
        count += 1
        z1 = nep.compute_Mlincomb(li, x.reshape(-1, 1))
        z2 = nep.compute_Mlincomb(li, x.reshape(-1, 1), np.array([1.0]), 1)
        dl = -np.vdot(y, z1) / np.vdot(y, z2)
        li += dl
    return li


def resinv(nep, lam=0.0, v=None, c=None, tol=np.finfo(float).eps * 100, maxit=100, linsolvercreator=None, errmeasure=None):
    n = nep.n
    lam = complex(lam)
    v = np.array(v, dtype=np.complex128)
    c = v.copy() if c is None else np.array(c, dtype=np.complex128)
    errmeasure = errmeasure or DefaultErrmeasure(nep)
    linsolver = (linsolvercreator or B200LinSolverCreator()).create_linsolver(nep, lam)
    use_v = np.linalg.norm(c) == 0
    err = np.inf
    for _ in range(maxit):
        v = v / np.linalg.norm(v)
        err = errmeasure.estimate_error(lam, v)
        if use_v:
            c = v.copy()
        if err < tol:
            return lam, v
        lam1 = compute_rf(nep, v, y=c, lam=lam)
        dv = -linsolver.lin_solve(nep.compute_Mlincomb(lam1, v.reshape(n, 1)))
        lam, v = lam1, v + dv
    raise NoConvergenceException(lam, v, err, "Number of iterations exceeded. maxit=%d." % maxit)


# ---------------------------------------------------------------------------------------------
# iar / tiar
# ---------------------------------------------------------------------------------------------
def iar(nep, maxit=30, linsolvercreator=None, tol=np.finfo(float).eps * 10000, neigs=6, errmeasure=None, sigma=0.0, gamma=1.0,
        v=None, check_error_every=1, orthmethod=dgks_host):
    n, m = nep.n, maxit
    sigma = complex(sigma)
    errmeasure = errmeasure or DefaultErrmeasure(nep)
    V = np.zeros((n * (m + 1), m + 1), dtype=np.complex128, order="F")
    H = np.zeros((m + 1, m), dtype=np.complex128)
    y = np.zeros((n, m + 1), dtype=np.complex128, order="F")
    alpha = np.asarray(gamma, dtype=np.complex128) ** np.arange(m + 1)
    alpha[0] = 0
    M0inv = (linsolvercreator or B200LinSolverCreator()).create_linsolver(nep, sigma)
    err = np.full((m, m), np.nan)
    lam = np.zeros(m + 1, dtype=np.complex128)
    Q = np.zeros((n, m + 1), dtype=np.complex128)
    v = np.random.default_rng(0).standard_normal(n) if v is None else np.asarray(v, dtype=np.complex128)
    V[:n, 0] = v / np.linalg.norm(v)
    k, conv_eig = 1, 0
    while k <= m and conv_eig < neigs:
        VV = V[:n * (k + 1), :k]
        vv = V[:n * (k + 1), k]
        y[:, 1:k + 1] = VV[:n * k, k - 1].reshape(n, k, order="F")
        y[:, 1:k + 1] /= np.arange(1, k + 1)[None, :]
        y[:, 0] = nep.compute_Mlincomb(sigma, y[:, :k + 1], alpha[:k + 1])
        y[:, 0] = -M0inv.lin_solve(y[:, 0].copy())
        vv[:] = y[:, :k + 1].reshape((k + 1) * n, order="F")
        H[k, k - 1] = orthmethod(VV, vv, H[:k, k - 1])
        if k % check_error_every == 0 or k == m:
            D, Zm = np.linalg.eig(H[:k, :k])
            Q = V[:n, :k] @ Zm
            lam = sigma + gamma / D
            e = _errs(errmeasure, lam, Q)
            err[k - 1, :len(lam)] = e
            conv_eig = int(np.count_nonzero(e < tol))
            idx = np.argsort(err[k - 1, :k], kind="stable")
            err[k - 1, :k] = err[k - 1, idx]
            if k == m or conv_eig >= neigs:
                nrof = int(min(len(lam), neigs))
                Q = Q[:, idx[:len(lam)]]
                lam = lam[idx[:nrof]]
        k += 1
    k -= 1
    if conv_eig < neigs and neigs != np.inf:
        raise NoConvergenceException(lam, Q, err[k - 1, :k], "Number of iterations exceeded. maxit=%d." % maxit)
    lam = lam[:min(len(lam), conv_eig)]
    Q = Q[:, :min(Q.shape[1], conv_eig)]
    return lam, Q, V[:, :k]


def tiar(nep, maxit=30, linsolvercreator=None, tol=np.finfo(float).eps * 10000, neigs=6, errmeasure=None, sigma=0.0, gamma=1.0,
         v=None, check_error_every=1, orthmethod=dgks_host):
    n, m = nep.n, maxit
    if n < m:
        raise LostOrthogonalityException("Loss of orthogonality in the matrix Z. The problem size is too small, use iar instead.")
    sigma = complex(sigma)
    errmeasure = errmeasure or DefaultErrmeasure(nep)
    a = np.zeros((m + 1, m + 1, m + 1), dtype=np.complex128)
    Z = np.zeros((n, m + 1), dtype=np.complex128, order="F")
    t = np.zeros(m + 1, dtype=np.complex128)
    H = np.zeros((m + 1, m), dtype=np.complex128)
    y = np.zeros((n, m + 1), dtype=np.complex128, order="F")
    alpha = np.asarray(gamma, dtype=np.complex128) ** np.arange(m + 1)
    alpha[0] = 0
    M0inv = (linsolvercreator or B200LinSolverCreator()).create_linsolver(nep, sigma)
    err = np.full((m + 1, m + 1), np.nan)
    lam = np.zeros(m + 1, dtype=np.complex128)
    Q = np.zeros((n, m + 1), dtype=np.complex128)
    v = np.random.default_rng(0).standard_normal(n) if v is None else np.asarray(v, dtype=np.complex128)
    Z[:, 0] = v / np.linalg.norm(v)
    a[0, 0, 0] = 1
    hist = np.zeros(m + 1, dtype=int)
    k, conv_eig = 1, 0
    while k <= m and conv_eig < neigs:
        y[:, 1:k + 1] = Z[:, :k] @ a[:k, k - 1, :k].T
        y[:, 1:k + 1] /= np.arange(1, k + 1)[None, :]
        y[:, 0] = nep.compute_Mlincomb(sigma, y[:, :k + 1], alpha[:k + 1])
        y[:, 0] = -M0inv.lin_solve(y[:, 0].copy())
        Z[:, k] = y[:, 0]
        t[k] = orthmethod(Z[:, :k], Z[:, k], t[:k])
        g = np.zeros((k + 1, k + 1), dtype=np.complex128)
        g[1:, :] = a[:k, k - 1, :k + 1] / np.arange(1, k + 1)[:, None]
        g[0, :] = t[:k + 1]
        h = np.einsum("ijl,il->j", a[:k, :k, :k].conj(), g[:k, :k])
        f = g
        f[:, :k] -= np.einsum("ijl,j->il", a[:k + 1, :k, :k], h)
        hh = np.einsum("ijl,il->j", a[:k, :k, :k].conj(), f[:k, :k])
        f[:, :k] -= np.einsum("ijl,j->il", a[:k + 1, :k, :k], hh)
        h = h + hh
        beta = np.linalg.norm(f)
        H[:k, k - 1] = h
        H[k, k - 1] = beta
        a[:k + 1, k, :k + 1] = f / beta
        if k % check_error_every == 0 or k == m:
            D, W = np.linalg.eig(H[:k, :k])
            VV = Z[:, :k] @ a[0, :k, :k].T
            Q = VV @ W
            lam = sigma + gamma / D
            e = _errs(errmeasure, lam, Q)
            err[k - 1, :len(lam)] = e
            conv_eig = int(np.count_nonzero(e < tol))
            idx = np.argsort(err[k - 1, :k], kind="stable")
            err[k - 1, :k] = err[k - 1, idx]
            if k == m or conv_eig >= neigs:
                nrof = int(min(len(lam), neigs))
                lam = lam[idx[:nrof]]
                Q = Q[:, idx[:nrof]]
            hist[k - 1] = conv_eig
        k += 1
    k -= 1
    if conv_eig < neigs and neigs != np.inf:
        raise NoConvergenceException(lam, Q, err[k - 1], "Number of iterations exceeded. maxit=%d." % maxit)
    lam = lam[:min(len(lam), conv_eig)]
    Q = Q[:, :min(Q.shape[1], conv_eig)]
    return lam, Q, Z[:, :k], hist


# ---------------------------------------------------------------------------------------------
# contour_block_SS (src/method_block_SS.jl:46-215, Shat_mode = :native): same quadrature seam, 2K moments
# ---------------------------------------------------------------------------------------------
def block_ss_quadrature(N, radius, sigma, K, rank=0, world=1):
    """Nodes / weights for the 2K moments Shat_j = (1/2 pi i) oint z^j M(z+sigma)^-1 V dz (method_block_SS.jl:135-152)."""
    radius = (radius, radius) if np.isscalar(radius) else tuple(radius)
    h = 2 * np.pi / N
    t = h * np.arange(N)
    g = radius[0] * np.cos(t) + 1j * radius[1] * np.sin(t)
    gp = -radius[0] * np.sin(t) + 1j * radius[1] * np.cos(t)
    W = np.stack([gp * g ** j * h / (2j * np.pi) for j in range(2 * K)], axis=1)
    mine = np.arange(rank, N, world)
    return mine, g[mine] + sigma, W[mine]


def block_ss_extract(Shat, U, sigma, K, rank_drop_tol):
    """method_block_SS.jl:156-214 (host: (K k) x (K k) Hankel matrices, SVD, generalised eigenproblem)."""
    import scipy.linalg as sla
    n, L, _ = Shat.shape
    Mhat = [U.conj().T @ Shat[:, :, j] for j in range(2 * K)]
    Hhat = np.block([[Mhat[i + j] for j in range(K)] for i in range(K)])
    Hhat2 = np.block([[Mhat[i + j + 1] for j in range(K)] for i in range(K)])
    UU, SS, VVh = np.linalg.svd(Hhat)
    mprime = int(np.count_nonzero(SS / SS[0] > rank_drop_tol))
    UU1, VV1 = UU[:, :mprime], VVh.conj().T[:, :mprime]
    xi, X = sla.eig(UU1.conj().T @ Hhat2 @ VV1, UU1.conj().T @ Hhat @ VV1)
    S = np.concatenate([Shat[:, :, j] for j in range(K)], axis=1)
    return sigma + xi, S @ VV1 @ X, mprime


def contour_block_SS(nep: B200SPMF, U=None, V=None, sigma=0.0, radius=1.0, N=1000, k=3, K=3, tol=np.sqrt(np.finfo(float).eps),
                     rank_drop_tol=None, batch=32, rank=0, world=1, return_moments=False, seed=10):
    n = nep.n
    rank_drop_tol = tol if rank_drop_tol is None else rank_drop_tol
    rng = np.random.default_rng(seed)
    U = rng.random((n, k)) if U is None else np.asarray(U)
    V = rng.random((n, k)) if V is None else np.asarray(V)
    k = V.shape[1]
    mine, lams, W = block_ss_quadrature(N, radius, sigma, K, rank, world)
    integ = ContourIntegrator(nep, k, 2 * K, min(batch, max(1, len(mine))))
    try:
        Shat, flags = integ.integrate(lams, W, V, reduce=world > 1)
    finally:
        integ.close()
    _check_node_flags(flags)
    lam, Vec, mprime = block_ss_extract(Shat, np.asarray(U, dtype=np.complex128), sigma, K, rank_drop_tol)
    if return_moments:
        return lam, Vec, Shat, mprime
    return lam, Vec


# ---------------------------------------------------------------------------------------------
# infbilanczos (src/method_infbilanczos.jl:33-244)
# ---------------------------------------------------------------------------------------------
def _taylor_hankel_blocks(nep, sigma, size):
    """Coefficient blocks of the bilinear form of infinite bi-Lanczos: C_t[i, j] = f_t^(i+j+1)(sigma) / (i+j+1)!, i, j < size,
    one size x size column-major block per SPMF term (GENERAL mode of the fused product)."""
    T = np.stack([f.taylor(sigma, 2 * size + 1) for f in nep.fi])
    idx = np.arange(size)[:, None] + np.arange(size)[None, :] + 1
    return np.stack([np.asfortranarray(T[t][idx]).reshape(-1, order="F") for t in range(len(nep.fi))])


def left_right_scalar_prod(nep, At, B, ma, mb, sigma):
    """left_right_scalar_prod (:227-244): sum_j At[:, j]' * (-sum_i M^(i+j-1)(sigma) B[:, i] / (i+j-1)!).  The reference runs
    the 'nasty double loop' of ma compute_Mlincomb calls (O(m^3 n) over the whole iteration); here it is ONE fused multi-term
    product Z = sum_t A_t (B C_t) with Hankel blocks of Taylor coefficients, followed by a dense inner product."""
    size = max(ma, mb)
    Z = nep.apply(_lib.COEF_GENERAL, B[:, :size], _taylor_hankel_blocks(nep, sigma, size), size)
    # columns of B beyond mb are zero by construction of the recurrences; columns of Z beyond ma are not used
    return -np.sum(np.conj(At[:, :ma]) * Z[:, :ma])


def infbilanczos(nep, nept, maxit=30, linsolvercreator=None, linsolvertcreator=None, v=None, u=None, tol=1e-12, neigs=5, errmeasure=None,
                 sigma=0.0, gamma=1, check_error_every=1):
    """Infinite bi-Lanczos for a device operator `nep` and its transposed problem `nept` (M(conj(lam))^H as its own operator
    with its own device LU, exactly as the reference takes it).  Host recurrences as in the reference (including its quirks: the
    left start vector is overwritten by the right one, :55, and gamma is unused); every O(n) product is a fused device SpMM.
    Returns (lam, Q, TT)."""
    n = nep.n
    sigma = complex(sigma)
    v = np.asarray(v, dtype=np.complex128)
    u = v.copy()
    errmeasure = errmeasure or DefaultErrmeasure(nep)
    M0inv = (linsolvercreator or B200LinSolverCreator()).create_linsolver(nep, sigma)
    M0Tinv = (linsolvertcreator or B200LinSolverCreator()).create_linsolver(nept, sigma)
    m = maxit

    def dmul(op, s, X):  # sum_i M^(i)(s) X[:, i-1] / i!  == compute_Mlincomb(op, s, X * Dk, ones, 1)
        k = X.shape[1]
        T = np.stack([f.taylor(s, k + 1)[1:] for f in op.fi])  # p x k: block t is the k-vector of Taylor coefficients 1..k
        return op.apply(_lib.COEF_GENERAL, X, T, 1)[:, 0]

    qt = M0Tinv.lin_solve(u)
    q = v / np.vdot(qt, dmul(nep, sigma, v.reshape(n, 1)))
    Z = lambda cols: np.zeros((n, cols), dtype=np.complex128, order="F")  # noqa: E731
    Q0, Qt0, Q1, Qt1 = Z(m + 1), Z(m + 1), Z(m + 1), Z(m + 1)
    R1, Rt1, R2, Rt2 = Z(m + 1), Z(m + 1), Z(m + 1), Z(m + 1)
    R1[:, 0], Rt1[:, 0] = q, qt
    Q_basis = Z(m + 1)
    alpha = np.zeros(m + 1, dtype=np.complex128)
    beta = np.zeros(m + 1, dtype=np.complex128)
    gam = np.zeros(m + 1, dtype=np.complex128)
    lam, Q, err = np.zeros(0, dtype=np.complex128), Z(0), np.zeros(0)
    for k in range(1, m + 1):
        omega = np.conj(left_right_scalar_prod(nep, Rt1, R1, k, k, sigma))
        beta[k - 1] = np.sqrt(abs(omega))
        gam[k - 1] = np.conj(omega) / beta[k - 1]
        Q1[:, :k] = R1[:, :k] / beta[k - 1]
        Qt1[:, :k] = Rt1[:, :k] / np.conj(gam[k - 1])
        Q_basis[:, k - 1] = Q1[:, 0]
        z2 = -M0inv.lin_solve(dmul(nep, sigma, Q1[:, :k]))
        zt2 = -M0Tinv.lin_solve(dmul(nept, np.conj(sigma), Qt1[:, :k]))
        R2[:, 0] = z2
        R2[:, 1:k + 1] = Q1[:, :k]
        Rt2[:, 0] = zt2
        Rt2[:, 1:k + 1] = Qt1[:, :k]
        if k > 1:
            R2[:, :k - 1] -= gam[k - 1] * Q0[:, :k - 1]
            Rt2[:, :k - 1] -= np.conj(beta[k - 1]) * Qt0[:, :k - 1]
        alpha[k] = left_right_scalar_prod(nep, Qt1, R2, k, k + 1, sigma)
        R2[:, :k] -= alpha[k] * Q1[:, :k]
        Rt2[:, :k] -= np.conj(alpha[k]) * Qt1[:, :k]
        R1, R2 = R2, R1
        R2[:] = 0
        Rt1, Rt2 = Rt2, Rt1
        Rt2[:] = 0
        Q0, Q1 = Q1, Q0
        Q1[:] = 0
        Qt0, Qt1 = Qt1, Qt0
        Qt1[:] = 0
        if k % check_error_every == 0 or k == m:
            omega = left_right_scalar_prod(nep, Rt1, R1, k + 1, k + 1, sigma)
            beta[k] = np.sqrt(abs(omega))
            gam[k] = np.conj(omega) / beta[k]
            TT = np.zeros((k + 1, k + 1), dtype=np.complex128)  # spdiagm(-1 => beta, 0 => alpha, 1 => gamma), k entries each
            i = np.arange(k)
            TT[i, i] = alpha[1:k + 1]
            TT[i + 1, i] = beta[1:k + 1]
            TT[i, i + 1] = gam[1:k + 1]
            with np.errstate(divide="ignore", invalid="ignore"):
                D, Zv = np.linalg.eig(TT)
                lam = sigma + 1.0 / D
            Q = Q_basis[:, :k + 1] @ Zv
            err = np.array([errmeasure.estimate_error(lam[s], Q[:, s]) if np.isfinite(lam[s]) else np.inf for s in range(len(lam))])
            conv_eig = int(np.count_nonzero(err < tol))
            idx = np.argsort(err[:k], kind="stable")
            err = err[idx]
            if conv_eig >= neigs or k == m:
                nrof = int(min(len(lam), neigs, conv_eig))
                lam = lam[idx[:nrof]]
                Q = Q[:, idx[:nrof]]
                if nrof:
                    Q = Q / np.linalg.norm(Q, axis=0)[None, :]
                if conv_eig >= neigs or neigs == np.inf:
                    return lam, Q, TT
    raise NoConvergenceException(lam, Q, err, "Number of iterations exceeded. maxit=%d." % maxit)


# ---------------------------------------------------------------------------------------------
# ilan (src/method_ilan.jl:56-261): infinite Lanczos for symmetric NEPs
# ---------------------------------------------------------------------------------------------
class _DenseLinSolverCreator:
    """Linear solves of the small dense projected problem (the reference's `factorize` of a dense matrix, LinSolvers.jl:116)."""

    class _Solver:
        def __init__(self, M):
            import scipy.linalg as L
            self._lu = L.lu_factor(np.asarray(M, dtype=np.complex128))
            self._solve = L.lu_solve

        def lin_solve(self, b, tol=0):
            return self._solve(self._lu, np.asarray(b, dtype=np.complex128))

    def create_linsolver(self, nep, lam):
        return self._Solver(nep.compute_Mder(lam))


def ilan_symmetrizer_coefficients(m):
    """symmetrizer_coefficients (method_ilan.jl:419-426)."""
    G = np.zeros((m + 1, m + 1))
    G[:, 0] = 1.0 / np.arange(1, m + 2)
    for j in range(1, m + 1):
        G[:, j] = G[:, j - 1] * j / (np.arange(1, m + 2) + j)
    return G


def ilan_inner_solve(pnep, neigs, tol=1e-13, maxit=80):
    """inner_solve(::IARInnerSolver, ...) on the projected problem (inner_solver.jl:308-346; the default for a projected SPMF,
    :249-250): iar from ones at sigma = 0; a NoConvergenceException still hands back what converged."""
    n = pnep.n
    try:
        resid = lambda l, w: float(np.linalg.norm(pnep.compute_Mlincomb(l, w)) / np.linalg.norm(w))  # ResidualErrmeasure(pnep)
        lam, V, _ = iar(pnep, sigma=0.0, neigs=neigs, tol=tol, maxit=maxit, v=np.ones(n), linsolvercreator=_DenseLinSolverCreator(),
                        errmeasure=resid)
        return lam, V
    except NoConvergenceException as e:
        return np.asarray(e.lam), np.asarray(e.v)


def ilan(nep: B200SPMF, maxit=30, linsolvercreator=None, tol=np.finfo(float).eps * 10000, neigs=6, errmeasure=None, sigma=0.0, gamma=1.0,
         v=None, check_error_every=30, proj_solve=True, inner_tol=1e-13, inner_maxit=80, orthmethod=dgks_host):
    """ilan with the SPMF B-multiplication (compute_Bmul_method_SPMF_NEP, method_ilan.jl:330-352,379-388) on the device operator.

    Per iteration the device does: compute_Mlincomb (GENERAL, q = 1), the shifted solve, and `Bmult!` -- the reference's loop of
    p dense products Qn (G .* FDH_t) followed by p sparse products (:383-387) is ONE fused multi-term product
    Z = sum_t A_t (Qn C_t) with C_t = (G .* FDH_t)[1:k+1, 1:k+1].  The three-term recurrence (bilinear forms without conjugation,
    `mat_sum`, :279-288), the DGKS step on the first blocks and the extraction (projected problem through B200ProjSPMF -- all
    A_t V in one fused pass -- solved with iar, or Ritz pairs of H) follow the reference on the host."""
    n, m = nep.n, maxit
    sigma, gamma = complex(sigma), complex(gamma)
    errmeasure = errmeasure or DefaultErrmeasure(nep)
    V = np.zeros((n, m + 1), dtype=np.complex128, order="F")
    Q = np.zeros((n, m + 1), dtype=np.complex128, order="F")
    Qp = np.zeros((n, m + 1), dtype=np.complex128, order="F")
    Qn = np.zeros((n, m + 1), dtype=np.complex128, order="F")
    H = np.zeros((m + 1, m), dtype=np.complex128)
    HH = np.zeros((m + 1, m), dtype=np.complex128)
    om = np.zeros(m + 1, dtype=np.complex128)
    a = gamma ** np.arange(2 * m + 3)
    a[0] = 0
    M0inv = (linsolvercreator or B200LinSolverCreator()).create_linsolver(nep, sigma)
    err = np.full((m, m), np.nan)
    W = np.zeros((n, m + 1), dtype=np.complex128)
    QQ = np.zeros((n, m + 1), dtype=np.complex128)
    # FDH_t[i, j] = fD[i + j, t] (1-based) = gamma^r f_t^(r)(sigma), r = i + j - 1: the first column of f_t(SS) for the bidiagonal SS
    G = ilan_symmetrizer_coefficients(m)
    r = np.arange(m + 1)[:, None] + np.arange(m + 1)[None, :] + 1
    fD = np.array([[gamma ** j * complex(f.derivative(sigma, j)) for j in range(2 * m + 2)] for f in nep.fi])  # p x (2m+2)
    GF = [G * fD[t][r] for t in range(nep.p)]
    v = np.random.default_rng(0).standard_normal(n) if v is None else np.asarray(v, dtype=np.complex128)
    Q[:, 0] = v / np.linalg.norm(v)
    om[0] = np.vdot(Q[:, 0], nep.compute_Mlincomb(0.0, np.stack([Q[:, 0], Q[:, 0]], axis=1), np.array([0.0, 1.0])))  # (:122: at 0)
    V[:, 0] = Q[:, 0]
    k, conv_eig = 1, 0
    lam = np.zeros(0, dtype=np.complex128)
    while k <= m and conv_eig < neigs:
        if not proj_solve:
            QQ[:, k - 1] = Q[:, 0]
        Qn[:, 1:k + 1] = Q[:, :k] / np.arange(1, k + 1)[None, :]
        Qn[:, 0] = nep.compute_Mlincomb(sigma, Qn[:, :k + 1], a[:k + 1])
        Qn[:, 0] = -M0inv.lin_solve(Qn[:, 0].copy())
        Cblk = np.stack([np.asfortranarray(GF[t][:k + 1, :k + 1]).reshape(-1, order="F") for t in range(nep.p)])
        Z = nep.apply(_lib.COEF_GENERAL, Qn[:, :k + 1], Cblk, k + 1)  # Bmult!
        beta = np.sum(Z[:, :k] * Qp[:, :k]) if k > 1 else 0.0
        alpha = np.sum(Z[:, :k] * Q[:, :k])
        eta = np.sum(Z[:, :k + 1] * Qn[:, :k + 1])
        H[k - 1, k - 1] = alpha / om[k - 1]
        if k > 1:
            H[k - 2, k - 1] = beta / om[k - 2]
        Qn[:, :k] -= H[k - 1, k - 1] * Q[:, :k]
        if k > 1:
            Qn[:, :k] -= H[k - 2, k - 1] * Qp[:, :k]
        H[k, k - 1] = np.linalg.norm(Qn)
        Qn[:, :k + 1] /= H[k, k - 1]
        om[k] = eta - 2 * alpha * H[k - 1, k - 1] + om[k - 1] * H[k - 1, k - 1] ** 2
        if k > 1:
            om[k] = om[k] - 2 * beta * H[k - 2, k - 1] + om[k - 2] * H[k - 2, k - 1] ** 2
        om[k] = om[k] / H[k, k - 1] ** 2
        V[:, k] = Qn[:, 0]
        HH[k, k - 1] = orthmethod(V[:, :k], V[:, k], HH[:k, k - 1])
        if k % check_error_every == 0 or k == m:
            if not proj_solve:
                D, Wr = np.linalg.eig(H[:k, :k])
                W[:, :k] = QQ[:, :k] @ Wr
                lam = sigma + gamma / D
            else:
                from .neptypes import create_proj_NEP
                VV = V[:, :k + 1]
                pnep = create_proj_NEP(nep, k + 1).set_projectmatrices(VV, VV)
                pnep.n = k + 1
                lproj, Wproj = ilan_inner_solve(pnep, m, inner_tol, inner_maxit)
                lam = np.asarray(lproj, dtype=np.complex128)
                q = min(len(lam), m)
                W[:, :q] = VV @ np.asarray(Wproj)[:, :q]
            nl = len(lam)
            err[k - 1, :nl] = _errs(errmeasure, lam, W[:, :nl]) if nl else []
            conv_eig = int(np.count_nonzero(err[k - 1, :nl] < tol))
            idx = np.argsort(err[k - 1, :k], kind="stable")
            err[k - 1, :k] = err[k - 1, idx]
            if k == m or conv_eig >= neigs:
                nrof = int(min(conv_eig, neigs))
                lam = lam[idx[:nrof]]
                W = W[:, idx[:len(lam)]]
        k += 1
        Qp[:] = Q
        Q[:] = Qn
        Qn[:] = 0
    k -= 1
    if conv_eig < neigs and neigs != np.inf:
        raise NoConvergenceException(lam, Q, err[k - 1, :k], "Number of iterations exceeded. maxit=%d." % maxit)
    return lam, W, err, V[:, :k + 1], H[:k, :k - 1], om[:k], HH[:k, :k]
