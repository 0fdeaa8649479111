"""Oracle: linear solves, contour integration and the solver loops that call the hot path
(test infrastructure, see oracle/__init__.py).

NumPy/SciPy restatement of
  * src/LinSolvers.jl:109-159 (FactorizeLinSolver / BackslashLinSolver) with SciPy's SuperLU standing in for
    UMFPACK (third-party in the reference: SuiteSparse_jll 5.10.1; parity pinned at the solution level only,
    test/linsolver.jl:21-68),
  * src/method_contour_common.jl:61-94 (integrate_interval, MatrixTrapezoidal),
  * src/method_beyncontour.jl:49-185 (contour_beyn),
  * src/method_iar.jl:47-184 (iar), src/method_tiar.jl:53-257 (tiar),
  * src/method_newton.jl:142-226 (resinv) with src/compute_rf_wrapper.jl:25-54 (scalar Newton Rayleigh functional),
  * IterativeSolvers 0.9.2 `orthogonalize_and_normalize!(V, w, h, DGKS)` (third-party, restated from its published
    algorithm: classical Gram-Schmidt, re-orthogonalise while ||w|| < ||h||/sqrt(2)).
The Beyn probe matrix is an argument: the reference draws it with Julia's randn after Random.seed!(10)
(method_beyncontour.jl:85-86), which cannot be reproduced outside Julia ("parity unpinned" for that draw).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from . import nep as o


class NoConvergenceException(Exception):
    def __init__(self, lam, v, errmeasure, msg):
        super().__init__(msg)
        self.lam, self.v, self.errmeasure = lam, v, errmeasure


class LostOrthogonalityException(Exception):
    pass


# --------------------------------------------------------------------------------------------
# linear solvers
# --------------------------------------------------------------------------------------------
class FactorizeLinSolver:
    def __init__(self, nep, lam):
        M = o.compute_Mder(nep, lam)
        if sp.issparse(M):
            self.lu = sla.splu(sp.csc_matrix(M, dtype=np.complex128))
            self.solve = self.lu.solve
        else:
            import scipy.linalg as L
            lu = L.lu_factor(np.asarray(M, dtype=np.complex128))
            self.solve = lambda b: L.lu_solve(lu, b)

    def lin_solve(self, b, tol=0):
        return self.solve(np.asarray(b, dtype=np.complex128))


class BackslashLinSolver:
    def __init__(self, nep, lam):
        self.nep, self.lam = nep, lam

    def lin_solve(self, b, tol=0):
        return FactorizeLinSolver(self.nep, self.lam).lin_solve(b)


class FactorizeLinSolverCreator:
    def create_linsolver(self, nep, lam):
        return FactorizeLinSolver(nep, lam)


class BackslashLinSolverCreator:
    def create_linsolver(self, nep, lam):
        return BackslashLinSolver(nep, lam)


# --------------------------------------------------------------------------------------------
# orthogonalisation (IterativeSolvers DGKS)
# --------------------------------------------------------------------------------------------
def orthogonalize_and_normalize_dgks(V, w, h):
    """In place on w and h; returns the norm.  V: (rows x k), w: rows, h: k."""
    h[:] = V.conj().T @ w
    w -= V @ h
    nrm = np.linalg.norm(w)
    eta = 1.0 / np.sqrt(2.0)
    projection_size = np.linalg.norm(h)
    while nrm < eta * projection_size:
        correction = V.conj().T @ w
        projection_size = np.linalg.norm(correction)
        w -= V @ correction
        h += correction
        nrm = np.linalg.norm(w)
    w *= 1.0 / nrm
    return nrm


# --------------------------------------------------------------------------------------------
# error measures
# --------------------------------------------------------------------------------------------
def default_errmeasure(nep):
    """errmeasure.jl:91-101: StandardSPMFErrmeasure for AbstractSPMF, otherwise the residual."""
    if isinstance(nep, (o.SPMF_NEP, o.PEP, o.DEP, o.SumNEP, o.DerSPMF)):
        return o.standard_spmf_errmeasure(nep)
    return o.residual_errmeasure(nep)


# --------------------------------------------------------------------------------------------
# contour integration
# --------------------------------------------------------------------------------------------
def integrate_interval_trapezoidal(f, gv, a, b, N):
    """method_contour_common.jl:61-94."""
    h = (b - a) / N
    t = a + h * np.arange(N)
    f1 = f(t[0])
    m = len(gv)
    S = np.zeros(f1.shape + (m,), dtype=np.complex128)
    G = np.zeros((N, m), dtype=np.complex128)
    for j in range(m):
        G[:, j] = [gv[j](tt) for tt in t]
    for i in range(N):
        temp = f1 if i == 0 else f(t[i])
        for j in range(m):
            S[:, :, j] += temp * G[i, j]
    return S * h


def beyn_extract(A0, A1, sigma, radius, k, neigs, tol, rank_drop_tol, errmeasure, sanity_check):
    """method_beyncontour.jl:113-184: SVD rank decision, B matrix, eigenpairs, filtering / sorting."""
    V, S, Wh = np.linalg.svd(A0, full_matrices=False)
    W = Wh.conj().T
    p = int(np.count_nonzero(S / S[0] > rank_drop_tol))
    V0, W0 = V[:, :p], W[:, :p]
    B = (V0.conj().T @ A1 @ W0) @ np.diag(1.0 / S[:p])
    lam, VB = np.linalg.eig(B)
    lam = lam + sigma
    Vv = V0 @ VB  # note: the reference's normalize! acts on a copy (:129-131), the columns stay as they are
    r1, r2 = radius
    if not sanity_check:
        sorted_index = np.argsort(np.abs(sigma - lam), kind="stable")
        ls = lam[sorted_index]
        inside = ((ls - sigma).real / r1) ** 2 + ((ls - sigma).imag / r2) ** 2 <= 1
        inside_perm = np.argsort(~inside, kind="stable")
        idx = sorted_index[inside_perm]
        return lam[idx], Vv[:, idx], {"p": p, "S": S}
    errs = np.array([errmeasure(lam[i], Vv[:, i]) for i in range(p)])
    good = np.nonzero(errs < tol)[0]
    sorted_good = good[np.argsort(np.abs(sigma - lam[good]), kind="stable")]
    ls = lam[sorted_good]
    inside = ((ls - sigma).real / r1) ** 2 + ((ls - sigma).imag / r2) ** 2 <= 1
    perm = np.argsort(~inside, kind="stable")
    idx = sorted_good[perm]
    if len(idx) > neigs:
        idx = idx[:neigs]
    return lam[idx], Vv[:, idx], {"p": p, "S": S, "errs": errs}


def contour_beyn(nep, Vh, sigma=0.0, radius=1.0, N=1000, neigs=2, k=None, tol=np.sqrt(np.finfo(float).eps),
                 linsolvercreator=None, errmeasure=None, sanity_check=True, rank_drop_tol=None, return_moments=False):
    """method_beyncontour.jl:49-185 with the probe matrix Vh (n x k) supplied by the caller."""
    n = nep.n
    k = neigs + 1 if k is None else k
    if k > n:
        raise ValueError("Cannot compute more eigenvalues than the size of the NEP with contour_beyn() k=%d n=%d" % (k, n))
    if k <= 0:
        raise ValueError("k must be positive, k=%d." % k)
    radius = (radius, radius) if np.isscalar(radius) else tuple(radius)
    rank_drop_tol = tol if rank_drop_tol is None else rank_drop_tol
    creator = linsolvercreator or BackslashLinSolverCreator()
    errmeasure = errmeasure or default_errmeasure(nep)
    Vh = np.asarray(Vh, dtype=np.complex128)
    g = lambda t: complex(radius[0] * np.cos(t), radius[1] * np.sin(t))
    gp = lambda t: complex(-radius[0] * np.sin(t), radius[1] * np.cos(t))

    def f(t):
        M0inv = creator.create_linsolver(nep, g(t) + sigma)
        return M0inv.lin_solve(Vh) * gp(t)

    AA = integrate_interval_trapezoidal(f, [lambda s: 1.0 + 0j, g], 0.0, 2 * np.pi, N)
    A0 = AA[:, :, 0] / (2j * np.pi)
    A1 = AA[:, :, 1] / (2j * np.pi)
    lam, V, info = beyn_extract(A0, A1, sigma, radius, k, neigs, tol, rank_drop_tol, errmeasure, sanity_check)
    if return_moments:
        return lam, V, A0, A1, info
    return lam, V


# --------------------------------------------------------------------------------------------
# resinv
# --------------------------------------------------------------------------------------------
def compute_rf_scalar_newton(nep, x, y=None, lam=0.0, tol=np.finfo(float).eps * 100, maxit=80):
    """compute_rf_wrapper.jl:25-54."""
    y = x if y is None else y
    lam_iter = complex(lam)
    dl = np.inf
    count = 0
    while abs(dl) > tol and count < maxit:
        count += 1
        z1 = o.compute_Mlincomb(nep, lam_iter, x.reshape(-1, 1))
        z2 = o.compute_Mlincomb(nep, lam_iter, x.reshape(-1, 1), np.array([1.0]), startder=1)
        dl = -np.vdot(y, z1) / np.vdot(y, z2)
        lam_iter += dl
    return lam_iter


def resinv(nep, lam=0.0, v=None, c=None, tol=np.finfo(float).eps * 100, maxit=100, linsolvercreator=None, errmeasure=None):
    """method_newton.jl:142-226 (armijo_factor = 1: no line search)."""
    n = nep.n
    lam = complex(lam)
    v = np.array(v, dtype=np.complex128)
    c = v.copy() if c is None else np.array(c, dtype=np.complex128)
    errmeasure = errmeasure or default_errmeasure(nep)
    linsolver = (linsolvercreator or FactorizeLinSolverCreator()).create_linsolver(nep, lam)
    use_v = np.linalg.norm(c) == 0
    err = np.inf
    for k in range(maxit):
        v = v / np.linalg.norm(v)
        err = errmeasure(lam, v)
        if use_v:
            c = v.copy()
        if err < tol:
            return lam, v
        lam1 = compute_rf_scalar_newton(nep, v, y=c, lam=lam)
        dlam = lam1 - lam
        dv = -linsolver.lin_solve(o.compute_Mlincomb(nep, lam1, v.reshape(n, 1)))
        lam = lam + dlam
        v = v + dv
    raise NoConvergenceException(lam, v, err, "Number of iterations exceeded. maxit=%d." % maxit)


# --------------------------------------------------------------------------------------------
# iar
# --------------------------------------------------------------------------------------------
def iar(nep, maxit=30, linsolvercreator=None, tol=np.finfo(float).eps * 10000, neigs=6, errmeasure=None, sigma=0.0, gamma=1.0,
        v=None, check_error_every=1, orth=orthogonalize_and_normalize_dgks):
    """method_iar.jl:47-184 without proj_solve."""
    n, m = nep.n, maxit
    sigma = complex(sigma)
    errmeasure = errmeasure or default_errmeasure(nep)
    V = np.zeros((n * (m + 1), m + 1), dtype=np.complex128, order="F")
    H = np.zeros((m + 1, m), dtype=np.complex128)
    y = np.zeros((n, m + 1), dtype=np.complex128, order="F")
    alpha = np.asarray(gamma, dtype=np.complex128) ** np.arange(m + 1)
    alpha[0] = 0
    M0inv = (linsolvercreator or FactorizeLinSolverCreator()).create_linsolver(nep, sigma)
    err = np.full((m, m), np.nan)
    lam = np.zeros(m + 1, dtype=np.complex128)
    Q = np.zeros((n, m + 1), dtype=np.complex128)
    v = np.asarray(v, dtype=np.complex128)
    V[:n, 0] = v / np.linalg.norm(v)
    k, conv_eig = 1, 0
    while k <= m and conv_eig < neigs:
        VV = V[:n * (k + 1), :k]
        vv = V[:n * (k + 1), k]
        y[:, 1:k + 1] = VV[:n * k, k - 1].reshape(n, k, order="F")
        y[:, 1:k + 1] /= np.arange(1, k + 1)[None, :]
        y[:, 0] = o.compute_Mlincomb(nep, sigma, y[:, :k + 1], alpha[:k + 1])
        y[:, 0] = -M0inv.lin_solve(y[:, 0].copy())
        vv[:] = y[:, :k + 1].reshape((k + 1) * n, order="F")
        H[k, k - 1] = orth(VV, vv, H[:k, k - 1])
        if k % check_error_every == 0 or k == m:
            D, Z = np.linalg.eig(H[:k, :k])
            Q = V[:n, :k] @ Z
            lam = sigma + gamma / D
            conv_eig = 0
            err[k - 1, :len(lam)] = [errmeasure(lam[s], Q[:, s]) for s in range(len(lam))]
            conv_eig = int(np.count_nonzero(err[k - 1, :len(lam)] < tol))
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


# --------------------------------------------------------------------------------------------
# tiar
# --------------------------------------------------------------------------------------------
def tiar(nep, maxit=30, linsolvercreator=None, tol=np.finfo(float).eps * 10000, neigs=6, errmeasure=None, sigma=0.0, gamma=1.0,
         v=None, check_error_every=1, orth=orthogonalize_and_normalize_dgks):
    """method_tiar.jl:53-257 without proj_solve.  Index translation: Julia a[i,j,l] == a[i-1,j-1,l-1] here."""
    n, m = nep.n, maxit
    if n < m:
        raise LostOrthogonalityException("Loss of orthogonality in the matrix Z. The problem size is too small, use iar instead.")
    sigma = complex(sigma)
    errmeasure = errmeasure or default_errmeasure(nep)
    a = np.zeros((m + 1, m + 1, m + 1), dtype=np.complex128)
    Z = np.zeros((n, m + 1), dtype=np.complex128, order="F")
    t = np.zeros(m + 1, dtype=np.complex128)
    H = np.zeros((m + 1, m), dtype=np.complex128)
    y = np.zeros((n, m + 1), dtype=np.complex128, order="F")
    alpha = np.asarray(gamma, dtype=np.complex128) ** np.arange(m + 1)
    alpha[0] = 0
    M0inv = (linsolvercreator or FactorizeLinSolverCreator()).create_linsolver(nep, sigma)
    err = np.full((m + 1, m + 1), np.nan)
    lam = np.zeros(m + 1, dtype=np.complex128)
    Q = np.zeros((n, m + 1), dtype=np.complex128)
    v = np.asarray(v, dtype=np.complex128)
    Z[:, 0] = v / np.linalg.norm(v)
    a[0, 0, 0] = 1
    conv_hist = np.zeros(m + 1, dtype=int)
    k, conv_eig = 1, 0
    while k <= m and conv_eig < neigs:
        y[:, 1:k + 1] = Z[:, :k] @ a[:k, k - 1, :k].T
        y[:, 1:k + 1] /= np.arange(1, k + 1)[None, :]
        y[:, 0] = o.compute_Mlincomb(nep, sigma, y[:, :k + 1], alpha[:k + 1])
        y[:, 0] = -M0inv.lin_solve(y[:, 0].copy())
        Z[:, k] = y[:, 0]
        t[k] = orth(Z[:, :k], Z[:, k], t[:k])
        g = np.zeros((k + 1, k + 1), dtype=np.complex128)
        for l in range(k + 1):
            g[1:k + 1, l] = a[:k, k - 1, l] / np.arange(1, k + 1)
            g[0, l] = t[l]
        h = np.zeros(k, dtype=np.complex128)
        for l in range(k):
            h += a[:k, :k, l].conj().T @ g[:k, l]
        f = g
        for l in range(k):
            f[:k + 1, l] -= a[:k + 1, :k, l] @ h
        hh = np.zeros(k, dtype=np.complex128)
        for l in range(k):
            hh += a[:k, :k, l].conj().T @ f[:k, l]
        for l in range(k):
            f[:k + 1, l] -= a[:k + 1, :k, l] @ hh
        h = h + hh
        beta = np.linalg.norm(f[:k + 1, :k + 1])
        H[:k, k - 1] = h
        H[k, k - 1] = beta
        a[:k + 1, k, :k + 1] = f[:k + 1, :k + 1] / beta
        if k % check_error_every == 0 or k == m:
            D, W = np.linalg.eig(H[:k, :k])
            VV = Z[:, :k] @ a[0, :k, :k].T
            Q = VV @ W
            lam = sigma + gamma / D
            err[k - 1, :len(lam)] = [errmeasure(lam[s], Q[:, s]) for s in range(len(lam))]
            conv_eig = int(np.count_nonzero(err[k - 1, :len(lam)] < tol))
            idx = np.argsort(err[k - 1, :k], kind="stable")
            err[k - 1, :k] = err[k - 1, idx]
            if k == m or conv_eig >= neigs:
                nrof = int(min(len(lam), neigs))
                lam = lam[idx[:nrof]]
                Q = Q[:, idx[:nrof]]
            conv_hist[k - 1] = conv_eig
        k += 1
    k -= 1
    if conv_eig < neigs and neigs != np.inf:
        raise NoConvergenceException(lam, Q, err[k - 1], "Number of iterations exceeded. maxit=%d." % maxit)
    lam = lam[:min(len(lam), conv_eig)]
    Q = Q[:, :min(Q.shape[1], conv_eig)]
    return lam, Q, Z[:, :k], conv_hist


# --------------------------------------------------------------------------------------------
# contour_block_SS (src/method_block_SS.jl:46-215, Shat_mode = :native)
# --------------------------------------------------------------------------------------------
def block_ss_extract(Shat, U, sigma, K, rank_drop_tol):
    """method_block_SS.jl:156-214: Hankel moment matrices, SVD rank cut, generalised eigenproblem, eigenvectors."""
    n, L, _ = Shat.shape
    Mhat = np.stack([U.conj().T @ Shat[:, :, j] for j in range(2 * K)], axis=2)
    m = K * L
    Hhat = np.zeros((m, m), dtype=np.complex128)
    Hhat2 = np.zeros((m, m), dtype=np.complex128)
    for i in range(K):
        for j in range(K):
            Hhat[i * L:(i + 1) * L, j * L:(j + 1) * L] = Mhat[:, :, i + j]
            Hhat2[i * L:(i + 1) * L, j * L:(j + 1) * L] = Mhat[:, :, i + j + 1]
    UU, SS, VVh = np.linalg.svd(Hhat)
    VV = VVh.conj().T
    mprime = int(np.count_nonzero(SS / SS[0] > rank_drop_tol))
    UU1, VV1 = UU[:, :mprime], VV[:, :mprime]
    import scipy.linalg as L_
    xi, X = L_.eig(UU1.conj().T @ Hhat2 @ VV1, UU1.conj().T @ Hhat @ VV1)
    S = np.concatenate([Shat[:, :, j] for j in range(K)], axis=1)
    return sigma + xi, S @ VV1 @ X, mprime


def contour_block_SS(nep, U, V, sigma=0.0, radius=1.0, N=1000, K=3, tol=np.sqrt(np.finfo(float).eps), linsolvercreator=None,
                     rank_drop_tol=None, return_moments=False):
    """U, V (n x k) are supplied by the caller (the reference draws them with Julia's rand after Random.seed!(10))."""
    radius = (radius, radius) if np.isscalar(radius) else tuple(radius)
    rank_drop_tol = tol if rank_drop_tol is None else rank_drop_tol
    creator = linsolvercreator or BackslashLinSolverCreator()
    V = np.asarray(V, dtype=np.complex128)
    g = lambda t: complex(radius[0] * np.cos(t), radius[1] * np.sin(t))
    gp = lambda t: complex(-radius[0] * np.sin(t), radius[1] * np.cos(t))

    def f(t):
        return creator.create_linsolver(nep, g(t) + sigma).lin_solve(V) * gp(t) / (2j * np.pi)

    gv = [(lambda s, kk=kk: g(s) ** kk) for kk in range(2 * K)]
    Shat = integrate_interval_trapezoidal(f, gv, 0.0, 2 * np.pi, N)
    lam, Vec, mprime = block_ss_extract(Shat, np.asarray(U, dtype=np.complex128), sigma, K, rank_drop_tol)
    if return_moments:
        return lam, Vec, Shat, mprime
    return lam, Vec


# --------------------------------------------------------------------------------------------
# iar_chebyshev (src/method_iar_chebyshev.jl)
# --------------------------------------------------------------------------------------------
def cheb_integration_matrix(m, a, b):
    """The hardcoded matrix L of method_iar_chebyshev.jl:130-131 (m x m)."""
    L = np.diag(np.concatenate([[2.0], 1.0 / np.arange(2, m + 1)]))
    if m > 2:
        L = L + np.diag(-1.0 / np.arange(1, m - 1), -2)
    return L * (b - a) / 4.0


def cheb_Tc(m, a, b):
    """T_i(c), i = 0..m, c = (a+b)/(a-b) (:238, :262, :274)."""
    cc = (a + b) / (a - b)
    return np.cos(np.arange(m + 1) * np.arccos(cc))


def DD0_mat_fun(f, S, sigma):
    """Divided-difference matrix function f[S, sigma I] from the block trick of :474-498."""
    n = S.shape[0]
    A = np.zeros((2 * n, 2 * n), dtype=np.complex128)
    A[:n, :n] = S
    A[:n, n:] = np.eye(n)
    A[n:, n:] = sigma * np.eye(n)
    return np.asarray(f(A), dtype=np.complex128)[:n, n:]


def cheb_precompute(nep, method, a, b, m, gamma, sigma):
    """precompute_data (:234-287) for the DEP, PEP and SPMF variants of compute_y0_cheb."""
    cc, kk = (a + b) / (a - b), 2.0 / (b - a)
    pre = {"Tc": cheb_Tc(m, a, b)}
    if method == "DEP":
        if sigma != 0 or gamma != 1:
            raise ValueError("This function does not support shift and scale parameters")
        Ttau = np.zeros((len(nep.tauv), m + 2))
        II = np.arange(m + 2)
        for j, tau in enumerate(nep.tauv):
            t = -kk * tau + cc
            if abs(t) <= 1:
                Ttau[j, :] = np.cos(II * np.arccos(t))
            elif t >= 1:
                Ttau[j, :] = np.cosh(II * np.arccosh(t))
            else:
                Ttau[j, :] = ((-1.0) ** II) * np.cosh(II * np.arccosh(-t))
        pre["Ttau"] = Ttau
    elif method == "PEP":
        if sigma != 0 or gamma != 1:
            raise ValueError("This function does not support shift and scale parameters")
        Li = np.linalg.inv(cheb_integration_matrix(m, a, b))
        pre["D"] = np.vstack([np.zeros((1, m)), Li[:m - 1, :]])
    elif method == "SPMF":
        Li = np.linalg.inv(cheb_integration_matrix(m, a, b))
        D = np.vstack([np.zeros((1, m)), Li[:m - 1, :]])
        DDs = sigma * np.eye(m) + gamma * D
        pre["DDf"] = [gamma * DD0_mat_fun(f, DDs, sigma) for f in o.get_fv(nep)]
    else:
        raise ValueError("unknown compute_y0 method " + str(method))
    return pre


def compute_y0_cheb(nep, method, X, Y, M0inv, pre):
    """compute_y0_cheb (:309-367)."""
    Tc = pre["Tc"]
    n, N = X.shape
    if method == "DEP":
        Av = o.get_Av(nep)
        y0 = X @ Tc[:N]
        for j in range(len(nep.tauv)):
            y0 = y0 - o._dot(Av[j + 1], Y @ pre["Ttau"][j, :N + 1])
        return M0inv.lin_solve(y0)
    if method == "PEP":
        d = len(nep.A) - 1
        v = Tc[:N].astype(np.complex128)
        y0 = np.zeros(n, dtype=np.complex128)
        for j in range(d):
            y0 = y0 + o._dot(nep.A[j + 1], X @ v)
            v = pre["D"][:N, :N] @ v
        y0 = -M0inv.lin_solve(y0)
        return y0 - Y @ Tc[:N + 1]
    Av = o.get_Av(nep)
    y0 = np.zeros((n, N), dtype=np.complex128)
    for i, A in enumerate(Av):
        y0 = y0 + o._dot(A, X @ pre["DDf"][i][:N, :N])
    y0 = y0 @ Tc[:N]
    y0 = -M0inv.lin_solve(y0)
    return y0 - Y @ Tc[:N + 1]


def iar_chebyshev(nep, maxit=30, linsolvercreator=None, tol=np.finfo(float).eps * 10000, neigs=6, errmeasure=None, sigma=0.0, gamma=1.0,
                  v=None, check_error_every=1, a=None, b=None, compute_y0_method="auto", orth=orthogonalize_and_normalize_dgks):
    """method_iar_chebyshev.jl:66-217.  compute_y0_method: "auto" (by type, :84-96), "DEP", "PEP" or "SPMF".  The DEP / PEP
    formulas need sigma = 0, gamma = 1 (the reference transforms the problem with shift_and_scale first, which is outside the
    restated path); the SPMF formula carries sigma and gamma itself.  Returns (lam, Q, err, V, H)."""
    method = compute_y0_method
    if a is None:  # :81-82: the delay interval for a DEP, [-1, 1] otherwise
        a = -float(np.max(nep.tauv)) if isinstance(nep, o.DEP) else -1.0
    if b is None:
        b = 0.0 if isinstance(nep, o.DEP) else 1.0
    if method == "auto":
        method = "DEP" if isinstance(nep, o.DEP) else "PEP" if isinstance(nep, o.PEP) else "SPMF"
    if method in ("DEP", "PEP") and (sigma != 0 or gamma != 1):
        raise NotImplementedError("shift_and_scale is not restated: use compute_y0_method='SPMF' with sigma / gamma")
    n, m = nep.n, maxit
    sigma = complex(sigma)
    errmeasure = errmeasure or default_errmeasure(nep)
    V = np.zeros((n * (m + 1), m + 1), dtype=np.complex128, order="F")
    H = np.zeros((m + 1, m), dtype=np.complex128)
    M0inv = (linsolvercreator or FactorizeLinSolverCreator()).create_linsolver(nep, sigma)
    err = np.ones((m, m))
    lam = np.zeros(m + 1, dtype=np.complex128)
    Q = np.zeros((n, m + 1), dtype=np.complex128)
    v = np.asarray(v, dtype=np.complex128)
    V[:n, 0] = v / np.linalg.norm(v)
    L = cheb_integration_matrix(m, a, b)
    pre = cheb_precompute(nep, method, a, b, m, gamma, sigma)
    k, conv_eig = 1, 0
    idx = None
    while k <= m and conv_eig < neigs:
        VV = V[:n * (k + 1), :k]
        vv = V[:n * (k + 1), k]
        X = VV[:n * k, k - 1].reshape(n, k, order="F")
        y = np.zeros((n, k + 1), dtype=np.complex128, order="F")
        y[:, 1:k + 1] = X @ L[:k, :k]
        y[:, 0] = compute_y0_cheb(nep, method, X, y, M0inv, pre)
        vv[:] = y.reshape((k + 1) * n, order="F")
        H[k, k - 1] = orth(VV, vv, H[:k, k - 1])
        if (k % check_error_every == 0 or k == m) and k > 2:
            D, Z = np.linalg.eig(H[:k, :k])
            Q = V[:n, :k] @ Z
            lam = sigma + gamma / D
            err[k - 1, :len(lam)] = [errmeasure(lam[s], Q[:, s]) for s in range(len(lam))]
            conv_eig = int(np.count_nonzero(err[k - 1, :len(lam)] < tol))
            idx = np.argsort(err[k - 1, :k], kind="stable")
            err[k - 1, :k] = err[k - 1, idx]
            if k == m or conv_eig >= neigs:
                nrof = int(min(len(lam), neigs))
                lam = lam[idx[:nrof]]
                Q = Q[:, idx[:nrof]]
        k += 1
    k -= 1
    if conv_eig < neigs and neigs != np.inf:
        msg = "Number of iterations exceeded. maxit=%d." % maxit
        if conv_eig < 3:
            msg += " Check that sigma is not an eigenvalue."
        raise NoConvergenceException(lam, Q, err[k - 1], msg)
    lam = lam[:min(len(lam), conv_eig)]
    Q = Q[:, :min(Q.shape[1], conv_eig)]
    return lam, Q, err[:k, :], V[:, :k], H[:k, :k]


# --------------------------------------------------------------------------------------------
# infbilanczos (src/method_infbilanczos.jl)
# --------------------------------------------------------------------------------------------
def left_right_scalar_prod(nep, nept, At, B, ma, mb, sigma):
    """method_infbilanczos.jl:227-244: sum_j At[:, j]' * (-sum_i M^(j+i-1)(sigma) B[:, i] / (j+i-1)!), the bilinear form of the
    infinite bi-Lanczos method (the double loop the reference calls 'nasty': O(m^3 n))."""
    c = 0.0
    for j in range(1, ma + 1):
        dd = 1.0 / np.array([math.factorial(t) for t in range(j, j + mb)], dtype=np.float64)
        XX = B[:, :mb] * dd[None, :]
        z = -o.compute_Mlincomb(nep, sigma, XX, np.ones(mb), j)
        c = c + np.vdot(At[:, j - 1], z)
    return c


def infbilanczos(nep, nept, maxit=30, linsolvercreator=None, linsolvertcreator=None, v=None, u=None, tol=1e-12, neigs=5, errmeasure=None,
                 sigma=0.0, gamma=1, check_error_every=1):
    """method_infbilanczos.jl:33-225.  `nept` is the transposed problem M(conj(lam))^H given as its own NEP (the reference takes
    it, and a second linear solver for it, as arguments).  As in the reference, the left starting vector is overwritten by the
    right one (`u=Vector{T}(v)`, :55) and gamma is unused.  Returns (lam, Q, TT)."""
    n = nep.n
    sigma = complex(sigma)
    v = np.asarray(v, dtype=np.complex128)
    u = v.copy()
    errmeasure = errmeasure or default_errmeasure(nep)
    M0inv = (linsolvercreator or FactorizeLinSolverCreator()).create_linsolver(nep, sigma)
    M0Tinv = (linsolvertcreator or FactorizeLinSolverCreator()).create_linsolver(nept, sigma)
    m = maxit
    qt = M0Tinv.lin_solve(u)
    q = v / np.vdot(qt, o.compute_Mlincomb(nep, sigma, v.reshape(n, 1), np.ones(1), 1))
    Z = lambda cols: np.zeros((n, cols), dtype=np.complex128)  # noqa: E731
    Q0, Qt0, Q1, Qt1 = Z(m), Z(m), Z(m), Z(m)
    R1, Rt1, R2, Rt2 = Z(m + 1), Z(m + 1), Z(m + 1), Z(m + 1)
    R1[:, 0], Rt1[:, 0] = q, qt
    Z2, Zt2 = Z(m), Z(m)
    Q_basis = Z(m + 1)
    alpha = np.zeros(m + 1, dtype=np.complex128)
    beta = np.zeros(m + 1, dtype=np.complex128)
    gam = np.zeros(m + 1, dtype=np.complex128)
    lam = np.zeros(0, dtype=np.complex128)
    Q = Z(0)
    err = np.zeros(0)
    TT = np.zeros((0, 0), dtype=np.complex128)
    for k in range(1, m + 1):
        omega = np.conj(left_right_scalar_prod(nep, nept, Rt1, R1, k, k, sigma))
        beta[k - 1] = np.sqrt(abs(omega))
        gam[k - 1] = np.conj(omega) / beta[k - 1]
        Q1[:, :k] = R1[:, :k] / beta[k - 1]
        Qt1[:, :k] = Rt1[:, :k] / np.conj(gam[k - 1])
        Q_basis[:, k - 1] = Q1[:, 0]
        Dk = 1.0 / np.array([math.factorial(t) for t in range(1, k + 1)], dtype=np.float64)
        Z2[:, k - 1] = -M0inv.lin_solve(o.compute_Mlincomb(nep, sigma, Q1[:, :k] * Dk[None, :], np.ones(k), 1))
        Zt2[:, k - 1] = -M0Tinv.lin_solve(o.compute_Mlincomb(nept, np.conj(sigma), Qt1[:, :k] * Dk[None, :], np.ones(k), 1))
        R2[:, 0] = Z2[:, k - 1]
        R2[:, 1:k + 1] = Q1[:, :k]
        if k > 1:
            R2[:, :k - 1] -= gam[k - 1] * Q0[:, :k - 1]
        Rt2[:, 0] = Zt2[:, k - 1]
        Rt2[:, 1:k + 1] = Qt1[:, :k]
        if k > 1:
            Rt2[:, :k - 1] -= np.conj(beta[k - 1]) * Qt0[:, :k - 1]
        alpha[k] = left_right_scalar_prod(nep, nept, Qt1, R2, k, k + 1, sigma)
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
            omega = left_right_scalar_prod(nep, nept, Rt1, R1, k + 1, k + 1, sigma)
            beta[k] = np.sqrt(abs(omega))
            gam[k] = np.conj(omega) / beta[k]
            a0, b0, g0 = alpha[1:k + 1], beta[1:k + 1], gam[1:k + 1]
            TT = np.zeros((k + 1, k + 1), dtype=np.complex128)  # spdiagm(-1 => b0, 0 => a0, 1 => g0): k entries on each diagonal
            idxk = np.arange(k)
            TT[idxk, idxk] = a0[:k]
            TT[idxk + 1, idxk] = b0[:k]
            TT[idxk, idxk + 1] = g0[:k]
            with np.errstate(divide="ignore", invalid="ignore"):
                D, Zv = np.linalg.eig(TT)
                lam = sigma + 1.0 / D
            Q = Q_basis[:, :k + 1] @ Zv
            err = np.array([errmeasure(lam[s], Q[:, s]) if np.isfinite(lam[s]) else np.inf for s in range(len(lam))], dtype=float)
            conv_eig = int(np.count_nonzero(err < tol))
            idx = np.argsort(err[:k], kind="stable")
            err = err[idx]
            if conv_eig >= neigs or k == m:
                nrof = int(min(len(lam), neigs, conv_eig))
                lam = lam[idx[:nrof]]
                Q = Q[:, idx[:nrof]]
                Q = Q / np.linalg.norm(Q, axis=0)[None, :] if nrof else Q
                if conv_eig >= neigs or neigs == np.inf:
                    return lam, Q, TT
    raise NoConvergenceException(lam, Q, err, "Number of iterations exceeded. maxit=%d." % maxit)


# --------------------------------------------------------------------------------------------
# ilan (src/method_ilan.jl): infinite Lanczos for symmetric NEPs, SPMF B-multiplication
# --------------------------------------------------------------------------------------------
def symmetrizer_coefficients(m):
    """method_ilan.jl:419-426."""
    G = np.zeros((m + 1, m + 1))
    G[:, 0] = 1.0 / np.arange(1, m + 2)
    for j in range(1, m + 1):
        for i in range(1, m + 2):
            G[i - 1, j] = G[i - 1, j - 1] * j / (i + j)
    return G


def ilan_precompute_spmf(nep, m, sigma, gamma):
    """method_ilan.jl:330-352: FDH[t][i,j] = fD[i+j, t] with fD[:, t] = f_t(SS)[:, 1], SS = sigma I + subdiag(gamma (1:2m+1))."""
    fv = o.get_fv(nep)
    SS = np.diag(sigma * np.ones(2 * m + 2, dtype=np.complex128)) + np.diag(gamma * np.arange(1, 2 * m + 2, dtype=np.complex128), -1)
    fD = np.stack([np.asarray(f(SS), dtype=np.complex128)[:, 0] for f in fv], axis=1)
    idx = np.arange(1, m + 2)[:, None] + np.arange(1, m + 2)[None, :]  # 1-based i + j, used as the 1-based row index of fD
    FDH = [fD[idx - 1, t] for t in range(len(fv))]
    return FDH, symmetrizer_coefficients(m)


def ilan_bmult_spmf(k, Qn, Av, FDH, G):
    """Bmult! for SPMF (method_ilan.jl:379-388): Z = sum_t A_t (Qn[:, 1:k+1] (G .* FDH_t)[1:k+1, 1:k+1])."""
    Z = np.zeros((Qn.shape[0], k + 1), dtype=np.complex128)
    for t, A in enumerate(Av):
        Z += o._dot(A, Qn[:, :k + 1] @ (G[:k + 1, :k + 1] * FDH[t][:k + 1, :k + 1]))
    return Z


def inner_solve_iar(pnep, neigs, tol=1e-13, maxit=80):
    """inner_solve(::IARInnerSolver, ...) (inner_solver.jl:308-346): iar on the projected problem from ones, sigma = 0; whatever
    converged is returned when the wanted number is not reached."""
    n = pnep.n
    try:
        lam, V, _ = iar(pnep, sigma=0.0, neigs=neigs, tol=tol, maxit=maxit, v=np.ones(n), errmeasure=o.residual_errmeasure(pnep))
        return lam, V
    except NoConvergenceException as e:
        return np.asarray(e.lam), np.asarray(e.v)


def ilan(nep, maxit=30, linsolvercreator=None, tol=np.finfo(float).eps * 10000, neigs=6, errmeasure=None, sigma=0.0, gamma=1.0, v=None,
         check_error_every=30, proj_solve=True, inner_tol=1e-13, inner_maxit=80, orth=orthogonalize_and_normalize_dgks):
    """method_ilan.jl:56-261 with the SPMF B-multiplication (compute_Bmul_method_SPMF_NEP; a DEP is taken through its SPMF
    form, which the reference's test "Different format" shows to be the same iteration)."""
    n, m = nep.n, maxit
    sigma, gamma = complex(sigma), complex(gamma)
    errmeasure = errmeasure or default_errmeasure(nep)
    V = np.zeros((n, m + 1), dtype=np.complex128)
    Q = np.zeros((n, m + 1), dtype=np.complex128)
    Qp = np.zeros((n, m + 1), dtype=np.complex128)
    Qn = np.zeros((n, m + 1), dtype=np.complex128)
    H = np.zeros((m + 1, m), dtype=np.complex128)
    HH = np.zeros((m + 1, m), dtype=np.complex128)
    om = np.zeros(m + 1, dtype=np.complex128)
    a = gamma ** np.arange(2 * m + 3)
    a[0] = 0
    M0inv = (linsolvercreator or FactorizeLinSolverCreator()).create_linsolver(nep, sigma)
    err = np.full((m, m), np.nan)
    W = np.zeros((n, m + 1), dtype=np.complex128)
    QQ = np.zeros((n, m + 1), dtype=np.complex128)
    FDH, G = ilan_precompute_spmf(nep, m, sigma, gamma)
    Av = o.get_Av(nep)
    v = np.asarray(v, dtype=np.complex128)
    Q[:, 0] = v / np.linalg.norm(v)
    om[0] = np.vdot(Q[:, 0], o.compute_Mlincomb(nep, 0, np.stack([Q[:, 0], Q[:, 0]], axis=1), np.array([0, 1])))  # (:122; at 0, not sigma)
    V[:, 0] = Q[:, 0]
    k, conv_eig = 1, 0
    lam = np.zeros(0, dtype=np.complex128)
    while k <= m and conv_eig < neigs:
        if not proj_solve:
            QQ[:, k - 1] = Q[:, 0]
        Qn[:, 1:k + 1] = Q[:, :k] / np.arange(1, k + 1)[None, :]
        Qn[:, 0] = o.compute_Mlincomb(nep, sigma, Qn[:, :k + 1], a[:k + 1])
        Qn[:, 0] = -M0inv.lin_solve(Qn[:, 0].copy())
        Z = ilan_bmult_spmf(k, Qn, Av, FDH, G)
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
        HH[k, k - 1] = orth(V[:, :k], V[:, k], HH[:k, k - 1])
        if k % check_error_every == 0 or k == m:
            if not proj_solve:
                D, Wr = np.linalg.eig(H[:k, :k])
                W[:, :k] = QQ[:, :k] @ Wr
                lam = sigma + gamma / D
            else:
                VV = V[:, :k + 1]
                pnep = o.create_proj_NEP(nep, k + 1)
                pnep.set_projectmatrices(VV, VV)
                lproj, Wproj = inner_solve_iar(pnep.nep_proj, m, inner_tol, inner_maxit)
                q = len(lproj)
                lam = np.asarray(lproj, dtype=np.complex128)
                q = min(q, m)
                W[:, :q] = VV @ np.asarray(Wproj)[:, :q]
            nl = len(lam)
            err[k - 1, :nl] = [errmeasure(lam[s], W[:, s]) for s in range(nl)]
            conv_eig = int(np.count_nonzero(err[k - 1, :nl] < tol))
            idx = np.argsort(err[k - 1, :k], kind="stable")  # NaNs (unused slots) sort last, as in Julia
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
