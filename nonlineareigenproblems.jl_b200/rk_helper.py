"""Rational-Krylov host helpers of nleigs (host-side mirror of src/rk_helper/*.jl).

Small O(#nodes) scalar work that stays on the host in the reference as well: discretisation of the target set
(discretizepolygon.jl), point-in-polygon test (inpolygon.jl), Leja-Bagby nodes / poles / scalings and scalar generalized
divided differences (rk_utils.jl), and the RKNEP classification (rk_nep.jl:101-126).  In the Julia deployment these are
NEP-PACK's own functions; nothing here touches the n-dimensional data.
"""
from __future__ import annotations

import warnings

import numpy as np
import scipy.linalg as sla

from .functions import Monomial
from .neptypes import PEP, SumNEP


def inpolygon(px, py, polyx, polyy):
    """Vectorised Hormann-Agathos test (inpolygon.jl:10-57): px, py arrays of points -> bool array; boundary = inside."""
    px = np.atleast_1d(np.asarray(px, dtype=np.float64))
    py = np.atleast_1d(np.asarray(py, dtype=np.float64))
    polyx = np.asarray(polyx, dtype=np.float64)
    polyy = np.asarray(polyy, dtype=np.float64)
    finite = np.isfinite(px) & np.isfinite(py)
    inside = np.zeros(px.shape, dtype=bool)  # crossing parity
    on = np.zeros(px.shape, dtype=bool)  # on a vertex or an edge
    m = len(polyx)
    px = np.where(finite, px, 0.0)  # non-finite points are rejected at the end; keep the arithmetic quiet
    py = np.where(finite, py, 0.0)
    for e in range(m):
        ax, ay = polyx[e], polyy[e]
        bx, by = polyx[(e + 1) % m], polyy[(e + 1) % m]
        on |= (ax == px) & (ay == py)
        lvl = by == py
        on |= lvl & (bx == px)
        on |= lvl & (ay == py) & ((bx > px) == (ax < px))
        cross = (ay < py) != (by < py)
        if not cross.any():
            continue
        det = (ax - px) * (by - py) - (bx - px) * (ay - py)
        need_det = cross & (((ax >= px) & ~(bx > px)) | (~(ax >= px) & (bx > px)))
        on |= need_det & (det == 0)
        flip = cross & (ax >= px) & (bx > px)
        flip |= need_det & (det != 0) & ((det > 0) == (by > ay))
        inside ^= flip
    return (inside | on) & finite


def in_sigma(z, Sigma, tol):
    """in_Σ (method_nleigs.jl:521-530); a real 2-point Σ is an interval thickened by tol."""
    Sigma = np.asarray(Sigma, dtype=np.complex128)
    z = np.atleast_1d(np.asarray(z, dtype=np.complex128))
    if len(Sigma) == 2 and not np.any(Sigma.imag):
        rx = np.array([Sigma[0].real, Sigma[0].real, Sigma[1].real, Sigma[1].real])
        ry = np.array([-tol, tol, tol, -tol])
    else:
        rx, ry = Sigma.real, Sigma.imag
    return inpolygon(z.real, z.imag, rx, ry)


def _walk_polygon(zc, npts):
    """npts points at equal arc length along the closed polygon zc (discretizepolygon.jl:38-58)."""
    L = np.abs(np.diff(zc)).sum()
    out = np.empty(npts, dtype=np.complex128)
    out[0] = zc[0]
    cnt, edge, alpha, rem = 1, 0, 0.0, L / npts
    while cnt < npts:
        d = abs(zc[edge + 1] - zc[edge])
        if (1 - alpha) * d < rem:
            rem -= (1 - alpha) * d
            edge, alpha = edge + 1, 0.0
        else:
            alpha += rem / d
            rem = L / npts
            out[cnt] = zc[edge] + alpha * (zc[edge + 1] - zc[edge])
            cnt += 1
    return out


def discretizepolygon(z, include_interior_points=False, npts=10000, nptsint=5):
    """discretizepolygon.jl:20-101: boundary points (then the vertices, closed) and optionally interior grid points.
    No vertex = unit disk, one = disk around it, two = Chebyshev points of the interval."""
    z = np.asarray(z, dtype=np.complex128).ravel()
    if z.size == 0:
        z = np.zeros(1, dtype=np.complex128)
    if z.size == 1:
        boundary = z[0] + np.exp(2j * np.pi * np.arange(1, npts + 1) / npts)
    elif z.size == 2:
        boundary = (z[1] - z[0]) / 2 * (np.cos(np.pi * np.arange(npts - 1, -1, -1) / (npts - 1)) + 1) + z[0]
    else:
        z = np.append(z, z[0])
        boundary = _walk_polygon(z, npts)
    zz = np.concatenate([boundary, z])
    interior = np.zeros(0, dtype=np.complex128)
    if not include_interior_points:
        return zz, interior
    if z.size == 2:
        cnt = 2 * nptsint + (1 if (2 * nptsint) % 2 == 0 else 0)
        return zz, np.linspace(z[0], z[1], cnt)[1::2].copy()
    pts = zz if z.size == 1 else z
    x0, x1, y0, y1 = pts.real.min(), pts.real.max(), pts.imag.min(), pts.imag.max()
    spacing = (x1 - x0) / 2.0001 / np.sqrt(nptsint)
    eps = np.finfo(float).eps
    for _ in range(10):
        nx, ny = int((x1 - x0) / (2 * spacing)), int((y1 - y0) / (2 * spacing))
        spacing /= 2.0 ** 0.25
        if nx <= 1 or ny <= 1:
            continue
        gx = np.linspace(x0, x1, nx)[1::2]
        gy = np.linspace(y0 - eps, y1 + eps, ny)[1::2]
        cand = (gx[:, None] + 1j * gy[None, :]).ravel()  # x-major, as the reference's comprehension
        interior = cand[inpolygon(cand.real, cand.imag, pts.real, pts.imag)]
        if len(interior) >= nptsint:
            return zz, interior
    raise RuntimeError("Failed to find interior polygon points. Polygon too narrow? (Note that intervals should be given by "
                       "their two endpoints only.)")


def lejabagby(A, B, Cset, m, keepA=False, forceInf=0):
    """rk_utils.jl:14-47: nodes a (greedy on A, or A itself), poles b (greedy on B, the first forceInf at infinity) and
    scalings beta normalising the nodal rational functions on Cset."""
    A = np.asarray(A, dtype=np.complex128)
    B = np.asarray(B, dtype=np.float64)
    Cset = np.asarray(Cset, dtype=np.complex128)
    if np.abs(B).min() < 1e-9:
        warnings.warn("There is at least one pole candidate in B being nearby zero. Consider shifting your problem for stability.")
    a = np.empty(m, dtype=np.complex128)
    b = np.empty(m, dtype=np.float64)
    beta = np.empty(m, dtype=np.float64)
    a[0], b[0], beta[0] = A[0], (np.inf if forceInf > 0 else B[0]), 1.0
    prodA, prodB, prodC = (np.ones(x.shape, dtype=np.complex128) for x in (A, B, Cset))
    with np.errstate(all="ignore"):
        for j in range(1, m):
            binv, scale = 1.0 / b[j - 1], 1.0 / beta[j - 1]
            prodA = prodA * scale * (A - a[j - 1]) / (1 - A * binv)
            prodB = prodB * scale * (B - a[j - 1]) / (1 - B * binv)
            prodC = prodC * scale * (Cset - a[j - 1]) / (1 - Cset * binv)
            if keepA:
                a[j] = A[j]
            else:
                mag = np.abs(prodA)
                mag[np.isnan(prodA)] = -np.inf
                a[j] = A[np.argmax(mag)]
            if forceInf > j:
                b[j] = np.inf
            else:
                mag = np.abs(prodB)
                mag[np.isnan(prodB)] = np.inf
                b[j] = B[np.argmin(mag)]
            beta[j] = np.abs(prodC).max()
            if beta[j] < np.finfo(float).eps:
                beta[j] = 1.0
    return a, b, beta


def evalrat(sigma, xi, beta, z):
    """rk_utils.jl:121-128."""
    z = np.asarray(z, dtype=np.complex128)
    r = np.full(z.shape, 1.0 / beta[0], dtype=np.complex128)
    for s, x, bt in zip(sigma, xi, beta[1:]):
        r = r * (z - s) / (1 - z / x) / bt
    return r


def ratnewtoncoeffs(fun, sigma, xi, beta):
    """rk_utils.jl:67-90 for a scalar function: divided differences by differencing (distinct nodes)."""
    m = len(sigma)
    d = np.zeros(m, dtype=np.complex128)
    d[0] = fun(complex(sigma[0])) * beta[0]
    for j in range(1, m):
        basis = np.array([evalrat(sigma[:k], xi[:k], beta[:k + 1], [sigma[j]])[0] for k in range(j + 1)])
        d[j] = (fun(complex(sigma[j])) - np.dot(d[:j], basis[:j])) / basis[j]
    return d


def ratnewtoncoeffsm(fm, sigma, xi, beta):
    """rk_utils.jl:96-118: all divided differences of a scalar function from the first column of fm(H K^-1), H and K the
    column-balanced bidiagonal pencil of the rational Newton basis."""
    m = len(sigma) - 1
    with np.errstate(all="ignore"):
        sub = np.asarray(beta[1:m + 1], dtype=np.float64) / np.asarray(xi[:m], dtype=np.float64)
    K = np.eye(m + 1, dtype=np.complex128)
    H = np.diag(np.asarray(sigma[:m + 1], dtype=np.complex128))
    idx = np.arange(m)
    K[idx + 1, idx] = sub
    H[idx + 1, idx] = beta[1:m + 1]
    scale = 1.0 / np.abs(K).max(axis=0)
    HK = sla.solve((K * scale).T, (H * scale).T).T
    return np.asarray(fm(HK), dtype=np.complex128)[:, 0] * beta[0]


def scgendivdiffs(sigma, xi, beta, maxdgr, isfunm, fv):
    """rk_utils.jl:57-67: sgdd[i, j] = j-th generalized divided difference of f_i."""
    out = np.zeros((len(fv), maxdgr + 2), dtype=np.complex128)
    for i, f in enumerate(fv):
        out[i, :] = ratnewtoncoeffsm(f, sigma, xi, beta) if isfunm else ratnewtoncoeffs(f, sigma, xi, beta)
    return out


def rk_structure(nep, source=None):
    """(p, q) of get_rk_nep (rk_nep.jl:101-126): PEP -> (d, 0); SumNEP(PEP, SPMF) -> (d, q); any other SPMF -> (-1, #terms).
    `source` is the host descriptor the device operator was built from (B200SPMF.from_nep keeps it)."""
    src = source if source is not None else getattr(nep, "source", None)
    nterms = len(nep.get_fv())
    if isinstance(src, PEP):
        return nterms - 1, 0
    if isinstance(src, SumNEP) and isinstance(src.nep1, PEP):
        return len(src.nep1.get_Av()) - 1, len(src.nep2.get_Av())
    if src is None and all(isinstance(f, Monomial) and f.d == i and f.c == 1.0 for i, f in enumerate(nep.get_fv())):
        return nterms - 1, 0  # an operator built directly from monomials is a PEP
    return -1, nterms


def low_rank_lu_factors(A):
    """low_rank_lu_factors + compactlu (rk_helper/rk_nep.jl:66-95): LU of the bounding box of the nonzeros of A with the trivial
    columns dropped, embedded back: A = L @ U.T with L (n x r), U (n x r) sparse.  As in the reference the row permutation of the
    dense LU is not carried along, so a factorisation that needs row exchanges is rejected."""
    import scipy.linalg as sl
    import scipy.sparse as sp
    A = sp.coo_matrix(A)
    n = A.shape[0]
    r0, r1, c0, c1 = A.row.min(), A.row.max(), A.col.min(), A.col.max()
    B = A.tocsr()[r0:r1 + 1, c0:c1 + 1].toarray()
    Pm, Lf, Uf = sl.lu(B)
    if not np.allclose(Pm, np.eye(len(B))):
        raise ValueError("low-rank LU factors with row exchanges are not supported (rk_nep.jl:80-88)")
    sel = np.array([(np.count_nonzero(Lf[i:, i]) > 1) or (np.count_nonzero(Uf[i, i:]) > 0) for i in range(len(B))])
    L = sp.lil_matrix((n, int(sel.sum())))
    L[r0:r1 + 1, :] = Lf[:, sel]
    U = sp.lil_matrix((A.shape[1], int(sel.sum())))
    U[c0:c1 + 1, :] = Uf[sel, :].T
    return sp.csr_matrix(L), sp.csr_matrix(U)
