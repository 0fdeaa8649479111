"""Oracle: the linear-algebra core of nleigs (test infrastructure, see oracle/__init__.py).

NumPy restatement of `backslash` (src/method_nleigs.jl:399-518) for the full-rank SPMF branch that large problems take
(`!P.is_low_rank`, `P.spmf && !computeD`, i.e. n > 400, :97-98,:456-462): the continuation vector wc holds N+1 blocks of
length n; B*wc is formed block by block, z0 collects -sum_i sgdd[i,ii+1] A_i z_ii through the stacked product
`P.BBCC * z_block` (src/rk_helper/rk_nep.jl:24,102-110 -- BBCC = vcat(A_1..A_p)), the first block is solved with the
shifted matrix from the solver cache, and the remaining blocks follow by substitution.
Indices below are 0-based: sigma[k] is the reference's sigma[k+1] (the current shift), xi[ii-1] its xi[ii], etc.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def backslash_fullrank(wc, Av, solve, sigma, k, beta, N, xi, sgdd):
    """Returns w = backslash(wc, ...).  `solve(shift, rhs)` is the cached shifted solve (linsolvercache.jl:7-26)."""
    n = Av[0].shape[0]
    BBCC = sp.vstack(Av).tocsr()  # rk_nep.jl:109
    shift = sigma[k]  # sigma[k+1] in the reference
    wc = np.asarray(wc, dtype=np.complex128)
    # construction of B*wc (:402-435)
    Bw = np.zeros_like(wc)
    for ii in range(1, N + 1):
        i0 = slice((ii - 1) * n, ii * n)
        i1 = slice(ii * n, (ii + 1) * n)
        Bw[i1] = wc[i0] + beta[ii] / xi[ii - 1] * wc[i1]
    # construction of z0 (:437-489)
    z = Bw.copy()
    nu = beta[1] * (1 - shift / xi[0])
    z[n:2 * n] = z[n:2 * n] / nu
    for ii in range(1, N + 1):
        i1 = slice(ii * n, (ii + 1) * n)
        prod = (BBCC @ z[i1]).reshape(-1, n).T  # reshape(P.BBCC * z_blk, n, :)
        z[:n] -= (prod * sgdd[:, ii][None, :]).sum(axis=1)
        if ii < N:
            i2 = slice((ii + 1) * n, (ii + 2) * n)
            mu = shift - sigma[ii]
            nu = beta[ii + 1] * (1 - shift / xi[ii])
            z[i2] = z[i2] / nu + mu / nu * z[i1]
    # solving Alam x0 = z0 (:491-494)
    w = np.zeros_like(wc)
    w[:n] = solve(shift, z[:n] / beta[0])
    # substitutions (:496-515)
    for ii in range(1, N + 1):
        i0 = slice((ii - 1) * n, ii * n)
        i1 = slice(ii * n, (ii + 1) * n)
        mu = shift - sigma[ii - 1]
        nu = beta[ii] * (1 - shift / xi[ii - 1])
        w[i1] = mu / nu * w[i0] + Bw[i1] / nu
    return w


# ================================================================================================
# The nleigs driver and its rational-Krylov helpers (src/method_nleigs.jl:60-377, src/rk_helper/*.jl)
# ================================================================================================
import warnings  # noqa: E402

import scipy.linalg as sla  # noqa: E402

from . import nep as o  # noqa: E402
from .solvers import orthogonalize_and_normalize_dgks, FactorizeLinSolverCreator  # noqa: E402


def det3p(q1x, q1y, q2x, q2y, px, py):
    """rk_helper/inpolygon.jl:59-61."""
    return (q1x - px) * (q2y - py) - (q2x - px) * (q1y - py)


def _isapprox0(det):
    # isapprox(0, det): |det| <= sqrt(eps)*max(0,|det|)  <=>  det == 0
    return det == 0


def inpolygon(px, py, polyx, polyy):
    """Hormann-Agathos point-in-polygon, points on the boundary count as inside (rk_helper/inpolygon.jl:10-57)."""
    if not (np.isfinite(px) and np.isfinite(py)):
        return False
    c = False
    m = len(polyx)
    for idx in range(m):
        q1x, q1y = polyx[idx], polyy[idx]
        q2x, q2y = polyx[(idx + 1) % m], polyy[(idx + 1) % m]
        if q1x == px and q1y == py:
            return True
        if q2y == py:
            if q2x == px:
                return True
            elif q1y == py and (q2x > px) == (q1x < px):
                return True
        if (q1y < py) != (q2y < py):
            if q1x >= px:
                if q2x > px:
                    c = not c
                else:
                    det = det3p(q1x, q1y, q2x, q2y, px, py)
                    if _isapprox0(det):
                        return True
                    elif (det > 0) == (q2y > q1y):
                        c = not c
            elif q2x > px:
                det = det3p(q1x, q1y, q2x, q2y, px, py)
                if _isapprox0(det):
                    return True
                elif (det > 0) == (q2y > q1y):
                    c = not c
    return c


def in_sigma(z, Sigma, tol):
    """method_nleigs.jl:521-530: membership of the points z in the polygonal target set."""
    Sigma = np.asarray(Sigma, dtype=np.complex128)
    if len(Sigma) == 2 and np.all(Sigma.imag == 0):
        rx = np.array([Sigma[0].real, Sigma[0].real, Sigma[1].real, Sigma[1].real])
        ry = np.array([-tol, tol, tol, -tol])
    else:
        rx, ry = Sigma.real, Sigma.imag
    return np.array([inpolygon(p.real, p.imag, rx, ry) for p in np.atleast_1d(z)], dtype=bool)


def discretizepolygon(z, include_interior_points=False, npts=10000, nptsint=5):
    """rk_helper/discretizepolygon.jl:20-101.  Returns (boundary points followed by the closed vertex list, interior)."""
    z = np.asarray(z, dtype=np.complex128).ravel()
    if len(z) == 0:
        z = np.zeros(1, dtype=np.complex128)
    if len(z) == 1:
        zz = z[0] + np.exp(1j * (2 * np.pi * np.arange(1, npts + 1) / npts))
    elif len(z) == 2:
        zz = (z[1] - z[0]) / 2 * (np.cos(np.pi * np.arange(npts - 1, -1, -1) / (npts - 1)) + 1) + z[0]
    else:
        z = np.concatenate([z, z[:1]])
        L = np.sum(np.abs(np.diff(z)))
        ind, alph = 0, 0.0
        pts = [z[0]]
        remL = L / npts
        while len(pts) < npts:
            d = abs(z[ind + 1] - z[ind])
            if (1 - alph) * d < remL:
                ind += 1
                remL -= (1 - alph) * d
                alph = 0.0
            else:
                alph += remL / d
                remL = L / npts
                pts.append(z[ind] + alph * (z[ind + 1] - z[ind]))
        zz = np.array(pts, dtype=np.complex128)
    zz = np.concatenate([zz, z])
    Z = np.zeros(0, dtype=np.complex128)
    if include_interior_points:
        if len(z) == 2:
            xnr = 2 * nptsint
            if xnr % 2 == 0:
                xnr += 1
            xpts = np.linspace(z[0], z[1], xnr)
            return zz, xpts[1::2].copy()
        points = zz if len(z) == 1 else z
        realz, imagz = points.real, points.imag
        real_min, real_max = realz.min(), realz.max()
        imag_min, imag_max = imagz.min(), imagz.max()
        it = 0
        spacing = (real_max - real_min) / 2.0001 / np.sqrt(nptsint)
        eps = np.finfo(float).eps
        while len(Z) < nptsint:
            it += 1
            if it > 10:
                raise RuntimeError("Failed to find interior polygon points. Polygon too narrow? (Note that intervals should be "
                                   "given by their two endpoints only.)")
            xnr = int((real_max - real_min) / (2 * spacing))
            ynr = int((imag_max - imag_min) / (2 * spacing))
            spacing /= np.sqrt(np.sqrt(2))
            if xnr <= 1 or ynr <= 1:
                continue
            xpts = np.linspace(real_min, real_max, xnr)[1::2]
            ypts = np.linspace(imag_min - eps, imag_max + eps, ynr)[1::2]
            cand = [complex(x, y) for x in xpts for y in ypts]
            Z = np.array([p for p in cand if inpolygon(p.real, p.imag, realz, imagz)], dtype=np.complex128)
    return zz, Z


def lejabagby(A, B, C, m, keepA=False, forceInf=0):
    """Greedy Leja-Bagby points a on A, poles b on B, scalings beta with unit uniform norm on C (rk_utils.jl:14-47)."""
    A = np.asarray(A, dtype=np.complex128)
    B = np.asarray(B, dtype=np.float64)
    C = np.asarray(C, dtype=np.complex128)
    if np.min(np.abs(B)) < 1e-9:
        warnings.warn("There is at least one pole candidate in B being nearby zero. Consider shifting your problem for stability.")
    a = [A[0]]
    b = [np.inf if forceInf > 0 else B[0]]
    beta = [1.0]
    sA = np.ones(A.shape, dtype=np.complex128)
    sB = np.ones(B.shape, dtype=np.complex128)
    sC = np.ones(C.shape, dtype=np.complex128)
    with np.errstate(all="ignore"):
        for j in range(m - 1):
            binv = 1.0 / b[j]
            binv_ = 1.0 / beta[j]
            sA = sA * binv_ * (A - a[j]) / (1 - A * binv)
            sB = sB * binv_ * (B - a[j]) / (1 - B * binv)
            sC = sC * binv_ * (C - a[j]) / (1 - C * binv)
            if keepA:
                a.append(A[j + 1])
            else:
                a.append(A[int(np.argmax(np.where(np.isnan(sA), -np.inf, np.abs(sA))))])
            if forceInf > j + 1:
                b.append(np.inf)
            else:
                b.append(B[int(np.argmin(np.where(np.isnan(sB), np.inf, np.abs(sB))))])
            bj = float(np.max(np.abs(sC)))
            if bj < np.finfo(float).eps:
                bj = 1.0
            beta.append(bj)
    return np.array(a, dtype=np.complex128), np.array(b, dtype=np.float64), np.array(beta, dtype=np.float64)


def evalrat(sigma, xi, beta, z):
    """Nodal rational function at the points z (rk_utils.jl:121-128)."""
    z = np.asarray(z, dtype=np.complex128)
    r = np.ones(z.shape, dtype=np.complex128) / beta[0]
    for j in range(len(sigma)):
        r = r * (z - sigma[j]) / (1 - z / xi[j]) / beta[j + 1]
    return r


def ratnewtoncoeffs(fun, sigma, xi, beta):
    """Rational divided differences by differencing; `fun` maps a 1x1 matrix to a (matrix) value (rk_utils.jl:67-90)."""
    m = len(sigma)
    D = [None] * m
    D[0] = np.atleast_2d(fun(np.array([[sigma[0]]], dtype=np.complex128))) * beta[0]
    for j in range(1, m):
        Qj = np.zeros(D[0].shape, dtype=np.complex128)
        for k in range(j):
            Qj = Qj + D[k] * evalrat(sigma[:k], xi[:k], beta[:k + 1], [sigma[j]])[0]
        D[j] = (np.atleast_2d(fun(np.array([[sigma[j]]], dtype=np.complex128))) - Qj) / evalrat(sigma[:j], xi[:j], beta[:j + 1], [sigma[j]])[0]
    return D


def ratnewtoncoeffsm(fm, sigma, xi, beta):
    """Rational divided differences of a scalar function through one matrix function (rk_utils.jl:96-118)."""
    m = len(sigma) - 1
    sigma = np.asarray(sigma, dtype=np.complex128)
    with np.errstate(all="ignore"):
        ksub = np.asarray(beta[1:m + 1], dtype=np.float64) / np.asarray(xi[:m], dtype=np.float64)
    K = np.diag(np.ones(m + 1, dtype=np.complex128)) + np.diag(ksub.astype(np.complex128), -1)
    H = np.diag(sigma[:m + 1]) + np.diag(np.asarray(beta[1:m + 1], dtype=np.complex128), -1)
    P = 1.0 / np.max(np.abs(K), axis=0)
    K = K * P[None, :]
    H = H * P[None, :]
    HK = sla.solve(K.T, H.T).T  # H / K
    D = np.asarray(fm(HK), dtype=np.complex128)[:, 0] * beta[0]
    return D


def scgendivdiffs(sigma, xi, beta, maxdgr, isfunm, pff):
    """Scalar generalized divided differences of every f_i (rk_utils.jl:57-67): sgdd[i, j]."""
    sgdd = np.zeros((len(pff), maxdgr + 2), dtype=np.complex128)
    for ii, f in enumerate(pff):
        if isfunm:
            sgdd[ii, :] = ratnewtoncoeffsm(f, sigma, xi, beta)
        else:
            sgdd[ii, :] = [np.asarray(d).ravel()[0] for d in ratnewtoncoeffs(lambda S: np.atleast_2d(f(S)), sigma, xi, beta)]
    return sgdd


class RKNEP:
    """get_rk_nep (rk_helper/rk_nep.jl:101-153): spmf flag, polynomial degree p, number of nonlinear terms q and, for a
    PEP + LowRankFactorizedNEP sum, the low-rank data (L factors side by side, UU = hcat(U...), iL = term of every column)."""

    def __init__(self, nep):
        self.nep = nep
        self.n = nep.n
        self.spmf = isinstance(nep, (o.SPMF_NEP, o.PEP, o.DEP, o.SumNEP, o.DerSPMF))
        self.p, self.q = 0, 0
        self.is_low_rank, self.r = False, 0
        self.Av = o.get_Av(nep) if self.spmf else []
        if not self.spmf:
            return
        if (isinstance(nep, o.SumNEP) and isinstance(nep.nep1, o.PEP) and isinstance(nep.nep2, o.LowRankFactorizedNEP)
                and len(o.get_Av(nep.nep2)) > 0):
            self.p, self.q = len(o.get_Av(nep.nep1)) - 1, len(o.get_Av(nep.nep2))
            self.is_low_rank, self.r = True, nep.nep2.r
            self.L = nep.nep2.L
            self.Lcat = sp.hstack(self.L).tocsr()          # n x r
            self.UU = sp.hstack(nep.nep2.U).tocsc()        # n x r
            self.iL = np.concatenate([np.full(L.shape[1], i) for i, L in enumerate(self.L)])
            return
        if isinstance(nep, o.PEP):
            self.p, self.q = len(self.Av) - 1, 0
        elif isinstance(nep, o.SumNEP) and isinstance(nep.nep1, o.PEP) and isinstance(nep.nep2, (o.SPMF_NEP, o.PEP, o.DEP)):
            self.p, self.q = len(o.get_Av(nep.nep1)) - 1, len(o.get_Av(nep.nep2))
        else:
            self.p, self.q = -1, len(self.Av)


class LinSolverCache:
    """rk_helper/linsolvercache.jl:7-26."""

    def __init__(self, nep, creator):
        self.solver, self.nep, self.creator = {}, nep, creator
        self.created = 0

    def solve(self, sigma, y, add_to_cache):
        key = complex(sigma)
        if add_to_cache:
            if key not in self.solver:
                self.solver[key] = self.creator.create_linsolver(self.nep, key)
                self.created += 1
            s = self.solver[key]
        else:
            s = self.creator.create_linsolver(self.nep, key)
            self.created += 1
        return s.lin_solve(y)


def _constructD(nb, P, sgdd):
    """method_nleigs.jl:380-396."""
    if P.is_low_rank and nb > P.p:
        return sp.hstack([sgdd[P.p + 1 + ii, nb] * P.L[ii] for ii in range(P.q)]).tocsr()
    D = None
    for ii, A in enumerate(P.Av):
        T = sgdd[ii, nb] * A
        D = T if D is None else D + T
    return D


def backslash_ref(wc, P, solve, computeD, sigma, k, D, beta, N, xi, sgdd):
    """The complete `backslash` of method_nleigs.jl:399-518 including the low-rank branches (`P.is_low_rank`): after the p-th
    block the blocks of the continuation vector have length r (sum of the ranks) instead of n, the step from block p to p + 1
    goes through UU^T, and the nonlinear terms enter z0 through the L factors (:463-471).  0-based: sigma[k] = the reference's
    sigma[k+1], beta[ii] = beta[ii+1], xi[ii-1] = xi[ii], D[ii] = D[ii+1], sgdd[:, ii] = sgdd[:, ii+1]."""
    n, p, lr = P.n, P.p, P.is_low_rank
    r = P.r if lr else 0
    shift = sigma[k]
    wc = np.asarray(wc, dtype=np.complex128)
    use_D = (not P.spmf) or computeD
    BBCC = None if use_D else sp.vstack(P.Av).tocsr()
    UUt = P.UU.conj().T.tocsr() if lr else None
    Bw = np.zeros_like(wc)
    if lr:  # first block (:408-416)
        blk = wc[(p - 1) * n:p * n]
        if use_D:
            Bw[:n] = -(D[p] @ blk) / beta[p]
        else:
            Bw[:n] = -((BBCC @ blk).reshape(-1, n).T * sgdd[:, p][None, :]).sum(axis=1) / beta[p]
    i0b, i0e = 0, n
    for ii in range(1, N + 1):  # other blocks (:418-435)
        i1b = i0e
        i1e = i0e + (n if (not lr or ii < p) else r)
        if not lr or ii != p:
            Bw[i1b:i1e] = wc[i0b:i0e] + beta[ii] / xi[ii - 1] * wc[i1b:i1e]
        else:
            Bw[i1b:i1e] = UUt @ wc[i0b:i0e] + beta[ii] / xi[ii - 1] * wc[i1b:i1e]
        i0b, i0e = i1b, i1e
    z = Bw.copy()  # construction of z0 (:437-489)
    i1b = n
    i1e = 2 * n if (not lr or p > 1) else n + r
    nu = beta[1] * (1 - shift / xi[0])
    z[i1b:i1e] = z[i1b:i1e] / nu
    for ii in range(1, N + 1):
        i2b = i1e
        i2e = i1e + (n if (not lr or ii < p - 1) else r)
        if use_D:
            if not lr or ii != p:
                z[:n] -= D[ii] @ z[i1b:i1e]
        else:
            if not lr or ii < p:
                z[:n] -= ((BBCC @ z[i1b:i1e]).reshape(-1, n).T * sgdd[:, ii][None, :]).sum(axis=1)
            elif ii > p:
                dd = sgdd[p + 1:, ii]
                z[:n] -= P.Lcat @ (z[i1b:i1e] * dd[P.iL])  # the LL / iLr loops of :463-470
        if ii < N:
            mu = shift - sigma[ii]
            nu = beta[ii + 1] * (1 - shift / xi[ii])
            if not lr or ii != p - 1:
                z[i2b:i2e] = z[i2b:i2e] / nu + mu / nu * z[i1b:i1e]
            else:
                z[i2b:i2e] = z[i2b:i2e] / nu + mu / nu * (UUt @ z[i1b:i1e])
        i1b, i1e = i2b, i2e
    w = np.zeros_like(wc)  # solve and substitutions (:491-515)
    w[:n] = solve(shift, z[:n] / beta[0])
    i0b, i0e = 0, n
    for ii in range(1, N + 1):
        i1b = i0e
        i1e = i0e + (n if (not lr or ii < p) else r)
        mu = shift - sigma[ii - 1]
        nu = beta[ii] * (1 - shift / xi[ii - 1])
        if not lr or ii != p:
            w[i1b:i1e] = mu / nu * w[i0b:i0e] + Bw[i1b:i1e] / nu
        else:
            w[i1b:i1e] = mu / nu * (UUt @ w[i0b:i0e]) + Bw[i1b:i1e] / nu
        i0b, i0e = i1b, i1e
    return w


def backslash_generic(wc, n, Dlist, solve, sigma, k, beta, N, xi, add_to_cache):
    """`backslash` with explicit divided-difference matrices D (computeD / non-SPMF branch, :456-459)."""
    shift = sigma[k]
    wc = np.asarray(wc, dtype=np.complex128)
    Bw = np.zeros_like(wc)
    for ii in range(1, N + 1):
        i0 = slice((ii - 1) * n, ii * n)
        i1 = slice(ii * n, (ii + 1) * n)
        Bw[i1] = wc[i0] + beta[ii] / xi[ii - 1] * wc[i1]
    z = Bw.copy()
    nu = beta[1] * (1 - shift / xi[0])
    z[n:2 * n] = z[n:2 * n] / nu
    for ii in range(1, N + 1):
        i1 = slice(ii * n, (ii + 1) * n)
        z[:n] -= Dlist[ii] @ z[i1]
        if ii < N:
            i2 = slice((ii + 1) * n, (ii + 2) * n)
            mu = shift - sigma[ii]
            nu = beta[ii + 1] * (1 - shift / xi[ii])
            z[i2] = z[i2] / nu + mu / nu * z[i1]
    w = np.zeros_like(wc)
    w[:n] = solve(shift, z[:n] / beta[0], add_to_cache)
    for ii in range(1, N + 1):
        i0 = slice((ii - 1) * n, ii * n)
        i1 = slice(ii * n, (ii + 1) * n)
        mu = shift - sigma[ii - 1]
        nu = beta[ii] * (1 - shift / xi[ii - 1])
        w[i1] = mu / nu * w[i0] + Bw[i1] / nu
    return w


def nleigs(nep, Sigma=(-1.0 - 1j, -1 + 1j, 1 + 1j, 1 - 1j), Xi=(np.inf,), maxdgr=100, minit=20, maxit=200, linsolvercreator=None, tol=1e-10,
           tollin=None, v=None, errmeasure=None, isfunm=True, static=False, leja=1, nodes=(), reusefact=1, blksize=20,
           return_details=False, check_error_every=5, backslash=None):
    """method_nleigs.jl:60-377, full-rank branches (SPMF with the stacked product, SPMF with explicit D for n <= 400, and
    the black-box NEP with matrix-valued divided differences).  0-based: sigma[k] here is the reference's sigma[k+1].
    `backslash` (tests only) replaces the full-rank SPMF backslash, e.g. by the device implementation.
    Returns (lam, X, res, details) with details a dict (Lam, Res, sigma, xi, beta, nrmD, kconv, iterations)."""
    Sigma = np.asarray(Sigma, dtype=np.complex128)
    Xi = np.asarray(Xi, dtype=np.float64)
    tollin = max(tol / 10, 100 * np.finfo(float).eps) if tollin is None else tollin
    P = RKNEP(nep)
    n = nep.n
    v = np.random.default_rng(0).standard_normal(n) if v is None else v
    v = np.asarray(v, dtype=np.complex128)
    errmeasure = errmeasure or o.residual_errmeasure(nep)
    nodes = np.asarray(nodes, dtype=np.complex128)
    if n == 1:
        maxdgr = maxit + 1
    computeD = n <= 400
    cache = LinSolverCache(nep, linsolvercreator or FactorizeLinSolverCreator())
    D = []
    # Discretization of Sigma --> Gamma & Leja-Bagby points (:120-146)
    if leja == 0:
        if len(nodes) == 0:
            raise ValueError("Interpolation nodes must be provided via 'nodes' when no Leja-Bagby points ('leja' == 0) are used.")
        gamma, _ = discretizepolygon(Sigma)
        max_count = maxit + maxdgr + 2 if static else max(maxit, maxdgr) + 2
        sigma = np.tile(nodes, int(np.ceil(max_count / len(nodes))))
        _, xi, beta = lejabagby(sigma[:maxdgr + 2], Xi, gamma, maxdgr + 2, True, P.p)
    elif leja == 1:
        if len(nodes) == 0:
            gamma, nodes = discretizepolygon(Sigma, True)
        else:
            gamma, _ = discretizepolygon(Sigma)
        nodes = np.tile(nodes, int(np.ceil((maxit + 1) / len(nodes))))
        sigma, xi, beta = lejabagby(gamma, Xi, gamma, maxdgr + 2, False, P.p)
    else:
        gamma, _ = discretizepolygon(Sigma)
        max_count = maxit + maxdgr + 2 if static else max(maxit, maxdgr) + 2
        sigma, xi, beta = lejabagby(gamma, Xi, gamma, max_count, False, P.p)
    xi = xi.copy()
    xi[maxdgr + 1] = np.nan
    sigma = np.array(sigma, dtype=np.complex128)
    if (not P.spmf or not isfunm) and len(sigma) != len(np.unique(sigma)):
        raise ValueError("All interpolation nodes must be distinct when no matrix functions are used for computing the "
                         "generalized divided differences.")
    # Rational Newton coefficients (:148-164)
    rng_ = slice(0, maxdgr + 2)
    if not P.spmf:
        D = ratnewtoncoeffs(lambda lam: _dense(o.compute_Mder(nep, lam[0, 0])), sigma[rng_], xi[rng_], beta[rng_])
        nrmD = [np.linalg.norm(D[0])]
        sgdd = np.zeros((0, 0), dtype=np.complex128)
    else:
        sgdd = scgendivdiffs(sigma[rng_], xi[rng_], beta[rng_], maxdgr, isfunm, o.get_fv(nep))
        if computeD:
            D.append(_constructD(0, P, sgdd))
        nrmD = [float(np.max(np.abs(sgdd[:, 0])))]
    if not np.isfinite(nrmD[0]):
        raise ValueError("The generalized divided differences must be finite.")

    # Rational Krylov (:166-359)
    kmax = maxit + maxdgr if static else maxit
    v = cache.solve(sigma[0], v / np.linalg.norm(v), reusefact == 2)
    Vrows = n * (2 if not static else 1)
    V = np.zeros((Vrows, min(blksize, kmax) + 1), dtype=np.complex128)
    V[:n, 0] = v / np.linalg.norm(v)
    H = np.zeros((kmax + 1, kmax), dtype=np.complex128)
    K = np.zeros((kmax + 1, kmax), dtype=np.complex128)
    Lam = np.zeros((kmax, kmax), dtype=np.complex128)
    Res = np.zeros((kmax, kmax))
    expand = True
    kconv = np.iinfo(np.int64).max // 2
    kn, l, N, nbconv, nblamin = n, 0, 0, 0, 0
    lam = np.zeros(0, dtype=np.complex128)
    X = np.zeros((n, 0), dtype=np.complex128)
    res = np.zeros(0)
    conv = np.zeros(0, dtype=bool)
    Av = P.Av

    def grow(rows, cols):
        nonlocal V
        if rows > V.shape[0] or cols > V.shape[1]:
            W = np.zeros((max(rows, V.shape[0]), max(cols, V.shape[1])), dtype=np.complex128)
            W[:V.shape[0], :V.shape[1]] = V
            V = W

    k = 1
    while k <= kmax:
        if expand:
            kn += n if (not P.is_low_rank or k < P.p) else P.r  # (:205-211)
            if P.spmf and computeD:
                D.append(_constructD(k, P, sgdd))
            N += 1
            if not P.spmf:
                nrmD.append(np.linalg.norm(D[k]))
            else:
                nrmD.append(float(np.max(np.abs(sgdd[:, k]))))  # out of bounds when the linearization never converges, as in the reference
            if not np.isfinite(nrmD[k]):
                raise ValueError("The generalized divided differences must be finite.")
            if n > 1 and k >= 5 and k < kconv:
                if sum(nrmD[k - 4:k + 1]) < 5 * tollin:
                    kconv = k - 1
                    if static:
                        kmax = maxit + kconv
                    expand = False
                    if leja == 1:
                        if len(sigma) < kmax + 1:
                            sigma = np.concatenate([sigma, np.zeros(kmax + 1 - len(sigma), dtype=np.complex128)])
                        sigma[k:kmax + 1] = nodes[:kmax - k + 1]
                    if not P.spmf or computeD:
                        D = D[:k]
                    xi, beta, nrmD = xi[:k], beta[:k], nrmD[:k]
                    if static:
                        kn -= n if (not P.is_low_rank or k < P.p) else P.r
                    N -= 1
                elif k == maxdgr + 1:
                    kconv = k
                    expand = False
                    if leja == 1:
                        if len(sigma) < kmax + 1:
                            sigma = np.concatenate([sigma, np.zeros(kmax + 1 - len(sigma), dtype=np.complex128)])
                        sigma[k:kmax + 1] = nodes[:kmax - k + 1]
                    N -= 1
                    warnings.warn("NLEIGS: Linearization not converged after %d iterations" % maxdgr)
        l = k - N if static else k
        if not static or (static and not expand):
            # V holds l + 1 vectors of kn rows after this step; columns are added a block at a time (:180-203)
            grow(kn, V.shape[1] if V.shape[1] >= l + 1 else min(kmax + 1, max(l + 1, V.shape[1] + blksize)))
            t = np.zeros(l, dtype=np.complex128)
            t[l - 1] = 1
            wc = V[:kn, l - 1].copy()
            add_to_cache = ((not expand or k > kconv) and reusefact == 1) or reusefact == 2
            if P.is_low_rank:
                w = backslash_ref(wc, P, lambda s, y: cache.solve(s, y, add_to_cache), computeD, sigma, k, D, beta, N, xi, sgdd)
            elif P.spmf and not computeD:
                if backslash is not None:
                    w = backslash(wc, sigma, k, beta, N, xi, sgdd, add_to_cache)
                else:
                    w = backslash_fullrank(wc, Av, lambda s, y: cache.solve(s, y, add_to_cache), sigma, k, beta, N, xi, sgdd)
            else:
                w = backslash_generic(wc, n, D, cache.solve, sigma, k, beta, N, xi, add_to_cache)
            h = np.zeros(l, dtype=np.complex128)
            H[l, l - 1] = orthogonalize_and_normalize_dgks(V[:kn, :l], w, h)
            H[:l, l - 1] = h
            K[:l, l - 1] = h * sigma[k] + t
            K[l, l - 1] = H[l, l - 1] * sigma[k]
            V[:kn, l] = w

        def check_convergence(all_):
            nonlocal lam, X, res, conv, nbconv, nblamin
            lambda_, S = sla.eig(K[:l, :l], H[:l, :l])
            if not all_:
                lamin = in_sigma(lambda_, Sigma, tol)
                ilam = np.nonzero(lamin)[0]
                lam = lambda_[ilam]
            else:
                ilam = np.nonzero(np.isfinite(lambda_))[0]
                lam = lambda_[ilam]
                lamin = in_sigma(lam, Sigma, tol)
            nblamin = int(np.sum(lamin))
            for i in ilam:
                S[:, i] = S[:, i] / np.linalg.norm(H[:l + 1, :l] @ S[:, i])
            X = V[:n, :l + 1] @ (H[:l + 1, :l] @ S[:, ilam])
            for i in range(X.shape[1]):
                X[:, i] = X[:, i] / np.linalg.norm(X[:, i])
            res = np.array([errmeasure(lam[i], X[:, i]) for i in range(len(lam))], dtype=float)
            conv = np.abs(res) < tol
            if all_:
                resall = np.full(l, np.nan)
                resall[ilam] = res
                si = sorted(range(l), key=lambda i: (abs(lambda_[i]), np.angle(lambda_[i])))
                Res[:l, l - 1] = resall[si]
                Lam[:l, l - 1] = lambda_[si]
                conv = conv & lamin
            nbconv = int(np.sum(conv)) if len(conv) else 0

        if not return_details and ((not expand and k >= N + minit and (k - (N + minit)) % check_error_every == 0) or
                                   (k >= kconv + minit and (k - (kconv + minit)) % check_error_every == 0) or k == kmax):
            check_convergence(False)
        elif return_details and (not static or (static and not expand)):
            check_convergence(True)
        if ((not expand and k >= N + minit) or k >= kconv + minit) and nblamin == nbconv:
            break
        k += 1

    details = {"Lam": Lam[:l, :l], "Res": Res[:l, :l], "sigma": sigma[:min(k, len(sigma))], "xi": xi[:k] if expand else xi,
               "beta": beta[:k] if expand else beta, "nrmD": nrmD[:k] if expand else nrmD, "kconv": kconv, "iterations": min(k, kmax),
               "factorizations": cache.created, "N": N, "l": l, "res_all": res, "lam_all": lam}
    if return_details and expand:
        warnings.warn("NLEIGS: Linearization not converged after %d iterations" % maxdgr)
    return lam[conv], X[:, conv], res[conv], details


def _dense(M):
    return M.toarray() if sp.issparse(M) else np.asarray(M)
