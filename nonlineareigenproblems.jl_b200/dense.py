"""Device-resident dense blocks for the infinite-Arnoldi callers (host-side mirror).

Mirrors the orthogonalisation hook `orthogonalize_and_normalize!(V, w, h, method)` that iar / tiar / nleigs dispatch on
(src/method_iar.jl:107, src/method_tiar.jl:128; a user subtype is the documented extension point, test/iar.jl:7-17) and
the tall-skinny products of src/method_tiar.jl:119,187-189 and src/method_iar.jl:114-115, with the Krylov basis kept in
HBM (nepb_block) so that no basis vector crosses PCIe inside the solver loop.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check, ptr
from .neptypes import Block, B200SPMF
from .linsolve import B200LinSolverCreator
from .solvers import DefaultErrmeasure, NoConvergenceException, LostOrthogonalityException


def dgks(V: Block, k: int, W: Block, wcol: int, rows: int = 0):
    """B200DGKS: returns (h, norm, sweeps); w = W[:, wcol] is orthogonalised against V[:, :k] and normalised in HBM."""
    h = np.zeros(max(k, 1), dtype=np.complex128)
    nrm, sweeps = C.c_double(), C.c_int()
    check(lib.nepb_orth_dgks(V._h, k, W._h, wcol, rows, ptr(h), C.byref(nrm), C.byref(sweeps)))
    return h[:k], nrm.value, sweeps.value


def block_gemm(A: Block, acol0: int, ka: int, Cm, Y: Block, ycol0: int, rows: int = 0):
    Cm = np.asfortranarray(np.asarray(Cm, dtype=np.complex128))
    if Cm.ndim == 1:
        Cm = Cm.reshape(-1, 1, order="F")
    assert Cm.shape[0] == ka
    check(lib.nepb_block_gemm(A._h, acol0, ka, ptr(Cm), Cm.shape[0], Cm.shape[1], Y._h, ycol0, rows))


def copy_cols(src: Block, s0: int, nc: int, dst: Block, d0: int, alpha=1.0, rows: int = 0):
    a = np.array([complex(alpha)], dtype=np.complex128)
    check(lib.nepb_block_copy_cols(src._h, s0, nc, dst._h, d0, ptr(a), rows))


def colnorms(A: Block, c0: int, nc: int, rows: int = 0):
    out = np.zeros(nc)
    check(lib.nepb_block_colnorms(A._h, c0, nc, rows, ptr(out)))
    return out


def solve_block(solver, Bb: Block, bcol0: int, nrhs: int, Xb: Block, xcol0: int, alpha=1.0, shift=0):
    """X[:, xcol0:xcol0+nrhs] = alpha * lin_solve(solver, B[:, bcol0:bcol0+nrhs]) with the operands in HBM.

    `solver` is whatever the linsolvercreator returned (the reference's extension point, LinSolvers.jl:100,124-137): a device
    factorisation (B200FactorizeLinSolver, or a bare B200LU) solves in place with the solver's `umfpack_refinements` steps of
    iterative refinement; any other LinSolver goes through its own `lin_solve` with host arrays."""
    lu = getattr(solver, "lu", solver)
    if hasattr(lu, "_h") and hasattr(lu, "nshift"):
        a = np.array([complex(alpha)], dtype=np.complex128)
        refine = int(getattr(solver, "refinements", 0))
        check(lib.nepb_lu_solve_block_ex(lu._h, shift, Bb._h, bcol0, nrhs, Xb._h, xcol0, ptr(a), refine, None))
        return
    B = Bb.download(bcol0, nrhs)
    X = np.asarray(solver.lin_solve(B if nrhs > 1 else B[:, 0]), dtype=np.complex128).reshape(Bb.n, nrhs, order="F")
    Xb.upload(alpha * X, xcol0)


def mlincomb_block(nep: B200SPMF, lam, Vb: Block, vcol0: int, k: int, a, Zb: Block, zcol0: int):
    """compute_Mlincomb!(nep, lam, V[:, vcol0:vcol0+k], a) -> Z[:, zcol0], operands in HBM."""
    if hasattr(nep, "mlincomb_block"):  # a NEP type with its own device product (WEP_FD, Waveguide.jl:324-379)
        return nep.mlincomb_block(lam, Vb, vcol0, k, a, Zb, zcol0)
    Cm, _ = nep.lincomb_coefficients(lam, np.asarray(a, dtype=np.complex128))
    Cm = np.ascontiguousarray(Cm.reshape(nep.p, k))
    check(lib.nepb_spmf_apply_block_ex(nep._h, _lib.COEF_GENERAL, Vb._h, vcol0, k, 1, ptr(Cm), Zb._h, zcol0))


def residual_errors(nep: B200SPMF, errmeasure, lams, Qb: Block, k: int, Rb: Block):
    """estimate_error for all k Ritz pairs at once: one multi-lambda SpMM + column norms, everything in HBM."""
    lams = np.asarray(lams, dtype=np.complex128)
    if hasattr(nep, "residual_block"):  # not an SPMF: one product per Ritz value (WEP_FD)
        return nep.residual_block(lams, Qb, k, Rb)
    Cd = np.empty((nep.p, k), dtype=np.complex128)
    for i, f in enumerate(nep.fi):
        Cd[i, :] = [complex(f(complex(s))) for s in lams]
    Cf = np.asfortranarray(Cd).reshape(-1, order="F")
    check(lib.nepb_spmf_apply_block_ex(nep._h, _lib.COEF_DIAG, Qb._h, 0, k, k, ptr(Cf), Rb._h, 0))
    r = colnorms(Rb, 0, k) / colnorms(Qb, 0, k)
    if hasattr(errmeasure, "_denom"):
        r = r / np.array([errmeasure._denom(l) for l in lams])
    return r


# ---------------------------------------------------------------------------------------------
# tiar with Z, y in HBM (src/method_tiar.jl:53-257)
# ---------------------------------------------------------------------------------------------
def tiar_device(nep: B200SPMF, maxit=30, linsolvercreator=None, tol=np.finfo(float).eps * 10000, neigs=6, errmeasure=None, sigma=0.0,
                gamma=1.0, v=None, check_error_every=1):
    n, m = nep.n, maxit
    if n < m:
        raise LostOrthogonalityException("Loss of orthogonality in the matrix Z. The problem size is too small, use iar instead.")
    sigma = complex(sigma)
    errmeasure = errmeasure or DefaultErrmeasure(nep)
    a = np.zeros((m + 1, m + 1, m + 1), dtype=np.complex128)
    t = np.zeros(m + 1, dtype=np.complex128)
    H = np.zeros((m + 1, m), dtype=np.complex128)
    alpha = np.asarray(gamma, dtype=np.complex128) ** np.arange(m + 1)
    alpha[0] = 0
    M0inv = (linsolvercreator or B200LinSolverCreator()).create_linsolver(nep, sigma)
    Zb, yb, tb = Block(n, m + 1), Block(n, m + 1), Block(n, 1)
    Qb, Rb = Block(n, m), Block(n, m)
    v = np.random.default_rng(0).standard_normal(n) if v is None else np.asarray(v, dtype=np.complex128)
    Zb.upload(v / np.linalg.norm(v), 0)
    a[0, 0, 0] = 1
    err = np.full((m + 1, m + 1), np.nan)
    lam = np.zeros(0, dtype=np.complex128)
    hist = np.zeros(m + 1, dtype=int)
    k, conv_eig = 1, 0
    idx = None
    while k <= m and conv_eig < neigs:
        # y[:, 1:k+1] = Z[:, :k] * a[:k, k-1, :k].T ./ (1:k)'   (tensor-core ZGEMM, column scaling folded into C)
        Cm = a[:k, k - 1, :k].T / np.arange(1, k + 1)[None, :]
        block_gemm(Zb, 0, k, Cm, yb, 1)
        mlincomb_block(nep, sigma, yb, 0, k + 1, alpha[:k + 1], tb, 0)
        solve_block(M0inv, tb, 0, 1, Zb, k, alpha=-1.0)  # Z[:, k] = -lin_solve(M0inv, y1)
        h0, t[k], _ = dgks(Zb, k, Zb, k)
        t[:k] = h0
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
            block_gemm(Zb, 0, k, a[0, :k, :k].T @ W, Qb, 0)  # Q = (Z a') W in one product
            lam = sigma + gamma / D
            e = residual_errors(nep, errmeasure, lam, Qb, k, Rb)
            err[k - 1, :k] = e
            conv_eig = int(np.count_nonzero(e < tol))
            idx = np.argsort(err[k - 1, :k], kind="stable")
            err[k - 1, :k] = err[k - 1, idx]
            hist[k - 1] = conv_eig
        k += 1
    k -= 1
    Q = Qb.download(0, len(lam)) if len(lam) else np.zeros((n, 0), complex)
    nrof = int(min(len(lam), neigs)) if idx is not None else 0
    if idx is not None:
        lam, Q = lam[idx[:nrof]], Q[:, idx[:nrof]]
    if conv_eig < neigs and neigs != np.inf:
        raise NoConvergenceException(lam, Q, err[k - 1], "Number of iterations exceeded. maxit=%d." % maxit)
    lam = lam[:min(len(lam), conv_eig)]
    Q = Q[:, :min(Q.shape[1], conv_eig)]
    return lam, Q, Zb.download(0, k), hist


# ---------------------------------------------------------------------------------------------
# iar with the n(m+1) x (m+1) basis in HBM (src/method_iar.jl:47-184)
# ---------------------------------------------------------------------------------------------
def iar_device(nep: B200SPMF, maxit=30, linsolvercreator=None, tol=np.finfo(float).eps * 10000, neigs=6, errmeasure=None, sigma=0.0,
               gamma=1.0, v=None, check_error_every=1, return_basis=True):
    n, m = nep.n, maxit
    sigma = complex(sigma)
    errmeasure = errmeasure or DefaultErrmeasure(nep)
    H = np.zeros((m + 1, m), dtype=np.complex128)
    alpha = np.asarray(gamma, dtype=np.complex128) ** np.arange(m + 1)
    alpha[0] = 0
    M0inv = (linsolvercreator or B200LinSolverCreator()).create_linsolver(nep, sigma)
    Vb = Block(n * (m + 1), m + 1)
    yb, tb = Block(n, m + 1), Block(n, 1)
    Qb, Rb = Block(n, m), Block(n, m)
    v = np.random.default_rng(0).standard_normal(n) if v is None else np.asarray(v, dtype=np.complex128)
    v0 = np.zeros(n * (m + 1), dtype=np.complex128)
    v0[:n] = v / np.linalg.norm(v)
    Vb.upload(v0, 0)
    err = np.full((m, m), np.nan)
    lam = np.zeros(0, dtype=np.complex128)
    idx = None
    k, conv_eig = 1, 0
    while k <= m and conv_eig < neigs:
        check(lib.nepb_iar_expand(Vb._h, k - 1, n, k, yb._h, 1, 1))  # y[:, 1:k+1] = reshape(VV[1:n*k, k], n, k) ./ (1:k)'
        mlincomb_block(nep, sigma, yb, 0, k + 1, alpha[:k + 1], tb, 0)
        solve_block(M0inv, tb, 0, 1, yb, 0, alpha=-1.0)
        check(lib.nepb_iar_pack(yb._h, 0, k + 1, n, Vb._h, k))  # vv = vec(y[:, 1:k+1])
        h, nrm, _ = dgks(Vb, k, Vb, k, rows=n * (k + 1))
        H[:k, k - 1] = h
        H[k, k - 1] = nrm
        if k % check_error_every == 0 or k == m:
            D, Zm = np.linalg.eig(H[:k, :k])
            block_gemm(Vb, 0, k, Zm, Qb, 0, rows=n)  # Q = V[1:n, 1:k] * Z
            lam = sigma + gamma / D
            e = residual_errors(nep, errmeasure, lam, Qb, k, Rb)
            err[k - 1, :k] = e
            conv_eig = int(np.count_nonzero(e < tol))
            idx = np.argsort(err[k - 1, :k], kind="stable")
            err[k - 1, :k] = err[k - 1, idx]
        k += 1
    k -= 1
    Q = Qb.download(0, len(lam)) if len(lam) else np.zeros((n, 0), complex)
    if idx is not None:
        nrof = int(min(len(lam), neigs))
        Q = Q[:, idx[:len(lam)]]
        lam = lam[idx[:nrof]]
    if conv_eig < neigs and neigs != np.inf:
        raise NoConvergenceException(lam, Q, err[k - 1, :k], "Number of iterations exceeded. maxit=%d." % maxit)
    lam = lam[:min(len(lam), conv_eig)]
    Q = Q[:, :min(Q.shape[1], conv_eig)]
    V = Vb.download(0, k) if return_basis else None
    return lam, Q, V


# ---------------------------------------------------------------------------------------------
# iar_chebyshev with the basis in HBM (src/method_iar_chebyshev.jl:66-217, compute_y0_cheb for AbstractSPMF :355-367)
# ---------------------------------------------------------------------------------------------
def cheb_integration_matrix(m, a, b):
    """The matrix L of method_iar_chebyshev.jl:130-131: coefficients of the antiderivative in the scaled Chebyshev basis."""
    L = np.diag(np.concatenate([[2.0], 1.0 / np.arange(2, m + 1)]))
    if m > 2:
        L = L + np.diag(-1.0 / np.arange(1, m - 1), -2)
    return L * (b - a) / 4.0


def cheb_divided_difference_blocks(nep: B200SPMF, m, a, b, gamma, sigma):
    """precompute_data for ComputeY0ChebSPMF_NEP (:270-287): DDf_i = gamma * f_i[sigma I + gamma D, sigma I] from the block
    matrix function f([[S, I], [0, sigma I]]) (:474-498); D = differentiation matrix in the Chebyshev basis."""
    Li = np.linalg.inv(cheb_integration_matrix(m, a, b))
    D = np.vstack([np.zeros((1, m)), Li[:m - 1, :]])
    S = sigma * np.eye(m) + gamma * D
    A = np.zeros((2 * m, 2 * m), dtype=np.complex128)
    A[:m, :m] = S
    A[:m, m:] = np.eye(m)
    A[m:, m:] = sigma * np.eye(m)
    return [gamma * np.asarray(f(A), dtype=np.complex128)[:m, m:] for f in nep.get_fv()]


def iar_chebyshev_device(nep: B200SPMF, maxit=30, linsolvercreator=None, tol=np.finfo(float).eps * 10000, neigs=6, errmeasure=None,
                         sigma=0.0, gamma=1.0, v=None, check_error_every=1, a=None, b=None):
    """iar_chebyshev for a device SPMF operator with the SPMF formula of compute_y0_cheb,
        y0 = -M(sigma)^-1 sum_i A_i X (DDf_i T(c)) - Y T(c),
    i.e. one fused multi-term product with k x 1 coefficient blocks, one device solve and two tall-skinny products per
    iteration; the n(m+1)-row basis, the DGKS sweeps and the Ritz residuals stay in HBM.  Returns (lam, Q, err, V, H)."""
    n, m = nep.n, maxit
    sigma = complex(sigma)
    src = getattr(nep, "source", None)
    tauv = getattr(src, "tauv", None)
    if a is None:  # :81-82: the delay interval for a DEP, [-1, 1] otherwise
        a = -float(np.max(tauv)) if tauv is not None else -1.0
    if b is None:
        b = 0.0 if tauv is not None else 1.0
    errmeasure = errmeasure or DefaultErrmeasure(nep)
    H = np.zeros((m + 1, m), dtype=np.complex128)
    M0inv = (linsolvercreator or B200LinSolverCreator()).create_linsolver(nep, sigma)
    L = cheb_integration_matrix(m, a, b)
    Tc = np.cos(np.arange(m + 1) * np.arccos((a + b) / (a - b)))
    DDf = cheb_divided_difference_blocks(nep, m, a, b, gamma, sigma)
    Vb = Block(n * (m + 1), m + 1)
    xb, yb, tb = Block(n, m), Block(n, m + 1), Block(n, 1)
    Qb, Rb = Block(n, m), Block(n, m)
    v = np.random.default_rng(0).standard_normal(n) if v is None else np.asarray(v, dtype=np.complex128)
    v0 = np.zeros(n * (m + 1), dtype=np.complex128)
    v0[:n] = v / np.linalg.norm(v)
    Vb.upload(v0, 0)
    err = np.ones((m, m))
    lam = np.zeros(0, dtype=np.complex128)
    idx = None
    nq = 0
    k, conv_eig = 1, 0
    while k <= m and conv_eig < neigs:
        check(lib.nepb_iar_expand(Vb._h, k - 1, n, k, xb._h, 0, 0))  # X = reshape(VV[1:n*k, k], n, k)
        block_gemm(xb, 0, k, L[:k, :k], yb, 1)                       # y[:, 2:k+1] = X * L[1:k, 1:k]
        Cm = np.ascontiguousarray(np.stack([Df[:k, :k] @ Tc[:k] for Df in DDf]))  # p x k: term i multiplies X by DDf_i T(c)
        check(lib.nepb_spmf_apply_block_ex(nep._h, _lib.COEF_GENERAL, xb._h, 0, k, 1, ptr(Cm), tb._h, 0))
        solve_block(M0inv, tb, 0, 1, yb, 0)                             # y[:, 1] = M0inv * (sum_i A_i X DDf_i T(c))
        coef = -np.concatenate([[1.0], Tc[1:k + 1]]).astype(np.complex128)
        block_gemm(yb, 0, k + 1, coef, tb, 0)                        # y0 = -y[:, 1] - y[:, 2:k+1] T(c)[2:k+1]
        copy_cols(tb, 0, 1, yb, 0)
        check(lib.nepb_iar_pack(yb._h, 0, k + 1, n, Vb._h, k))
        h, nrm, _ = dgks(Vb, k, Vb, k, rows=n * (k + 1))
        H[:k, k - 1] = h
        H[k, k - 1] = nrm
        if (k % check_error_every == 0 or k == m) and k > 2:
            D, Zm = np.linalg.eig(H[:k, :k])
            block_gemm(Vb, 0, k, Zm, Qb, 0, rows=n)
            nq = k
            lam = sigma + gamma / D
            e = residual_errors(nep, errmeasure, lam, Qb, k, Rb)
            err[k - 1, :k] = e
            conv_eig = int(np.count_nonzero(e < tol))
            idx = np.argsort(err[k - 1, :k], kind="stable")
            err[k - 1, :k] = err[k - 1, idx]
        k += 1
    k -= 1
    Q = Qb.download(0, nq) if nq else np.zeros((n, 0), complex)
    if idx is not None:
        nrof = int(min(len(lam), neigs))
        lam, Q = lam[idx[:nrof]], Q[:, idx[:nrof]]
    if conv_eig < neigs and neigs != np.inf:
        msg = "Number of iterations exceeded. maxit=%d." % maxit
        if conv_eig < 3:
            msg += " Check that sigma is not an eigenvalue."
        raise NoConvergenceException(lam, Q, err[k - 1], msg)
    lam = lam[:min(len(lam), conv_eig)]
    Q = Q[:, :min(Q.shape[1], conv_eig)]
    V = Vb.download(0, k)
    for blk in (Vb, xb, yb, tb, Qb, Rb):
        blk.close()
    return lam, Q, err[:k, :], V, H[:k, :k]
