"""nleigs' linear-algebra core on the device (host-side mirror of `backslash`, src/method_nleigs.jl:399-518, full-rank
SPMF branch).

The reference walks the N blocks of the continuation vector one after the other: N stacked SpMVs `BBCC*z_ii` (each reads
all p matrices, :462), 3N block axpys and one cached shifted solve.  The block recurrences do not involve the first
block, so on the device the whole routine is
    Bw = Wc * CB,  Z = Bw * CZ            two tall-skinny products (n x (N+1)) * ((N+1) x (N+1)) on the FP64 tensor cores
    z0 = -sum_i A_i (Z[:, 1:N] c_i)       ONE fused multi-term SpMM, c_i = sgdd[i, 1:N]  (GENERAL mode, q = 1)
    w0 = M(shift)^-1 z0 / beta_0          device LU from the solver cache
    W  = [w0, Bw[:, 1:N]] * CW            one more tall-skinny product
with the small coefficient matrices CB, CZ, CW built on the host from sigma, xi, beta exactly as the reference's scalars.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import lib, check, ptr
from .neptypes import B200SPMF, Block, SumNEP, PEP, LowRankFactorizedNEP
from .dense import block_gemm, solve_block
from .linsolve import B200FactorizeLinSolver


class DeviceLinSolverCache:
    """LinSolverCache (rk_helper/linsolvercache.jl:7-26) holding device factorisations keyed by the shift."""

    def __init__(self, nep: B200SPMF, umfpack_refinements=0):
        self.nep = nep
        self.solvers = {}
        self.refinements = umfpack_refinements

    def get(self, shift, add_to_cache=True):
        key = complex(shift)
        s = self.solvers.get(key)
        if s is None:
            s = B200FactorizeLinSolver(self.nep, shift, self.refinements)
            if add_to_cache:
                self.solvers[key] = s
        return s


def backslash_coefficients(sigma, k, beta, N, xi):
    """CB, CZ, CW ((N+1) x (N+1), column j = coefficients of output block j) and the shift."""
    shift = sigma[k]
    m = N + 1
    CB = np.zeros((m, m), dtype=np.complex128)
    for ii in range(1, N + 1):  # Bw_ii = wc_{ii-1} + beta_ii/xi_{ii-1} wc_ii
        CB[ii - 1, ii] = 1.0
        CB[ii, ii] = beta[ii] / xi[ii - 1]
    CZ = np.zeros((m, m), dtype=np.complex128)  # z_j as a combination of the Bw blocks
    nu = beta[1] * (1 - shift / xi[0])
    CZ[1, 1] = 1.0 / nu
    for ii in range(1, N):
        mu = shift - sigma[ii]
        nu = beta[ii + 1] * (1 - shift / xi[ii])
        CZ[:, ii + 1] = mu / nu * CZ[:, ii]
        CZ[ii + 1, ii + 1] += 1.0 / nu
    CW = np.zeros((m, m), dtype=np.complex128)  # w_j from [w0, Bw_1..Bw_N]
    CW[0, 0] = 1.0
    for ii in range(1, N + 1):
        mu = shift - sigma[ii - 1]
        nu = beta[ii] * (1 - shift / xi[ii - 1])
        CW[:, ii] = mu / nu * CW[:, ii - 1]
        CW[ii, ii] += 1.0 / nu
    return shift, CB, CZ, CW


def nleigs_backslash(nep: B200SPMF, cache: DeviceLinSolverCache, wc, sigma, k, beta, N, xi, sgdd, add_to_cache=True):
    """w = backslash(wc, ...) with wc, w host vectors of length n*(N+1); all O(n) work runs on the device."""
    n = nep.n
    m = N + 1
    shift, CB, CZ, CW = backslash_coefficients(sigma, k, beta, N, xi)
    Wc = np.asarray(wc, dtype=np.complex128).reshape(n, m, order="F")
    wcb, bwb, zb, t = Block.from_host(Wc), Block(n, m), Block(n, m), Block(n, 1)
    block_gemm(wcb, 0, m, CB, bwb, 0)
    block_gemm(bwb, 0, m, CZ, zb, 0)
    # z0 = Bw_0 (= 0) - sum_i A_i (Z[:, 1:N] sgdd[i, 1:N]); one fused pass over all terms
    Cm = -np.ascontiguousarray(np.asarray(sgdd, dtype=np.complex128)[:, 1:N + 1])  # p x N: block i = N-vector
    check(lib.nepb_spmf_apply_block_ex(nep._h, _lib.COEF_GENERAL, zb._h, 1, N, 1, ptr(Cm), t._h, 0))
    solver = cache.get(shift, add_to_cache)
    solve_block(solver, t, 0, 1, bwb, 0, alpha=1.0 / beta[0])  # Bw_0 is unused from here on: it receives w0
    block_gemm(bwb, 0, m, CW, wcb, 0)
    w = wcb.download()
    for b in (wcb, bwb, zb, t):
        b.close()
    return w.reshape(n * m, order="F")


# =================================================================================================
# The nleigs driver with the rational Krylov basis in HBM (src/method_nleigs.jl:60-377)
# =================================================================================================
import warnings  # noqa: E402

import scipy.linalg as sla  # noqa: E402

from . import rk_helper as rk  # noqa: E402
from .dense import dgks, copy_cols, residual_errors  # noqa: E402
from .solvers import ResidualErrmeasure  # noqa: E402


class _Basis:
    """V of the reference (kn x (l+1), kn growing by n per expansion step) as one row-major device block of
    n*blocks rows and `cols` columns: growing the row count appends memory, earlier vectors stay zero-padded."""

    def __init__(self, n, cols, blocks=4):
        self.n, self.cols, self.blocks = n, cols, blocks
        self.b = Block(n * blocks, cols)

    def ensure(self, blocks, used_cols):
        if blocks <= self.blocks:
            return
        nb = max(blocks, min(2 * self.blocks, self.blocks + 32))
        new = Block(self.n * nb, self.cols)
        if used_cols > 0:
            copy_cols(self.b, 0, used_cols, new, 0, rows=self.n * self.blocks)
        self.b.close()
        self.b, self.blocks = new, nb

    def close(self):
        self.b.close()


def _release(solver):
    """Free the device factors of a solver that owns some (user-supplied LinSolvers may not)."""
    lu = getattr(solver, "lu", None)
    if lu is not None and hasattr(lu, "close"):
        lu.close()


def _device_backslash(nep, cache, basis, l, wcb, bwb, zb, tb, sigma, k, beta, N, xi, sgdd, add_to_cache):
    """backslash (:399-518) for continuation vector V[:, l-1]; the result is packed into column l of the basis.
    Same products as nleigs_backslash, operands resident in HBM."""
    n, m = nep.n, N + 1
    shift, CB, CZ, CW = backslash_coefficients(sigma, k, beta, N, xi)
    check(lib.nepb_iar_expand(basis.b._h, l - 1, n, m, wcb._h, 0, 0))
    block_gemm(wcb, 0, m, CB, bwb, 0)
    block_gemm(bwb, 0, m, CZ, zb, 0)
    Cm = -np.ascontiguousarray(np.asarray(sgdd, dtype=np.complex128)[:, 1:N + 1])
    check(lib.nepb_spmf_apply_block_ex(nep._h, _lib.COEF_GENERAL, zb._h, 1, N, 1, ptr(Cm), tb._h, 0))
    solver = cache.get(shift, add_to_cache)
    solve_block(solver, tb, 0, 1, bwb, 0, alpha=1.0 / beta[0])
    block_gemm(bwb, 0, m, CW, wcb, 0)
    check(lib.nepb_iar_pack(wcb._h, 0, m, n, basis.b._h, l))
    if not add_to_cache and cache.solvers.get(complex(shift)) is not solver:
        _release(solver)  # a one-off factorisation (linsolvercache.jl:21-23); never a cached one


def nleigs(nep: B200SPMF, Sigma=(-1.0 - 1j, -1 + 1j, 1 + 1j, 1 - 1j), Xi=(np.inf,), maxdgr=100, minit=20, maxit=200, tol=1e-10,
           tollin=None, v=None, errmeasure=None, isfunm=True, static=False, leja=1, nodes=(), reusefact=1, blksize=20,
           return_details=False, check_error_every=5, umfpack_refinements=0, poly_degree=None):
    """nleigs for a device SPMF operator: the reference's control flow (:100-377, full-rank SPMF branch `P.spmf &&
    !computeD`) on the host, every O(n) operation on the device -- the stacked product of `backslash` as one fused SpMM,
    block recurrences as DMMA products, the cached shifted solves on the device LU, DGKS on the n(N+1)-row basis in HBM,
    Ritz vectors and residuals of all candidates in one product each.  `errmeasure`: None = ResidualErrmeasure
    (the reference default), a device error measure, or a host callable (lam, x) -> float as in the reference's gun tests.
    Returns (lam, X, res, details)."""
    src = getattr(nep, "source", None)
    if (poly_degree is None and isinstance(src, SumNEP) and isinstance(src.nep1, PEP) and isinstance(src.nep2, LowRankFactorizedNEP)
            and len(src.nep2.get_Av()) > 0):  # get_rk_nep's low-rank case (rk_helper/rk_nep.jl:127-152)
        return nleigs_lowrank(nep, LowRankStructure(len(src.nep1.get_Av()) - 1, src.nep2.L, src.nep2.U), Sigma, Xi, maxdgr, minit,
                              maxit, tol, tollin, v, errmeasure, isfunm, static, leja, nodes, reusefact, check_error_every,
                              umfpack_refinements)
    Sigma = np.asarray(Sigma, dtype=np.complex128)
    Xi = np.asarray(Xi, dtype=np.float64)
    n = nep.n
    tollin = max(tol / 10, 100 * np.finfo(float).eps) if tollin is None else tollin
    p_poly = rk.rk_structure(nep)[0] if poly_degree is None else poly_degree
    v = np.random.default_rng(0).standard_normal(n) if v is None else v
    v = np.asarray(v, dtype=np.complex128)
    errmeasure = errmeasure or ResidualErrmeasure(nep)
    host_err = callable(errmeasure) and not hasattr(errmeasure, "estimate_error")
    nodes = np.asarray(nodes, dtype=np.complex128)
    if n == 1:
        maxdgr = maxit + 1
    # -- nodes, poles, scalings (:120-146) -------------------------------------------------------------------------
    if leja == 0:
        if len(nodes) == 0:
            raise ValueError("Interpolation nodes must be provided via 'nodes' when no Leja-Bagby points ('leja' == 0) are used.")
        gamma, _ = rk.discretizepolygon(Sigma)
        max_count = maxit + maxdgr + 2 if static else max(maxit, maxdgr) + 2
        sigma = np.tile(nodes, -(-max_count // len(nodes)))
        _, xi, beta = rk.lejabagby(sigma[:maxdgr + 2], Xi, gamma, maxdgr + 2, True, p_poly)
    elif leja == 1:
        if len(nodes) == 0:
            gamma, nodes = rk.discretizepolygon(Sigma, True)
        else:
            gamma, _ = rk.discretizepolygon(Sigma)
        nodes = np.tile(nodes, -(-(maxit + 1) // len(nodes)))
        sigma, xi, beta = rk.lejabagby(gamma, Xi, gamma, maxdgr + 2, False, p_poly)
    else:
        gamma, _ = rk.discretizepolygon(Sigma)
        max_count = maxit + maxdgr + 2 if static else max(maxit, maxdgr) + 2
        sigma, xi, beta = rk.lejabagby(gamma, Xi, gamma, max_count, False, p_poly)
    xi[maxdgr + 1] = np.nan
    if not isfunm and len(sigma) != len(np.unique(sigma)):
        raise ValueError("All interpolation nodes must be distinct when no matrix functions are used for computing the "
                         "generalized divided differences.")
    head = slice(0, maxdgr + 2)
    sgdd = rk.scgendivdiffs(sigma[head], xi[head], beta[head], maxdgr, isfunm, nep.get_fv())
    nrmD = [float(np.abs(sgdd[:, 0]).max())]
    if not np.isfinite(nrmD[0]):
        raise ValueError("The generalized divided differences must be finite.")

    # -- rational Krylov (:166-359) ----------------------------------------------------------------------------------
    kmax = maxit + maxdgr if static else maxit
    cache = DeviceLinSolverCache(nep, umfpack_refinements)
    first = cache.get(sigma[0], reusefact == 2)
    v = first.lin_solve(v / np.linalg.norm(v))
    if cache.solvers.get(complex(sigma[0])) is not first:
        _release(first)
    cols = kmax + 1
    basis = _Basis(n, cols)
    col0 = np.zeros(n * basis.blocks, dtype=np.complex128)
    col0[:n] = v / np.linalg.norm(v)
    basis.b.upload(col0, 0)
    H = np.zeros((kmax + 1, kmax), dtype=np.complex128)
    K = np.zeros((kmax + 1, kmax), dtype=np.complex128)
    Lam = np.zeros((kmax, kmax), dtype=np.complex128)
    Res = np.zeros((kmax, kmax))
    work = {"m": 0}
    Qb, Rb = Block(n, cols), Block(n, cols)
    tb = Block(n, 1)
    expand, kconv = True, np.iinfo(np.int64).max // 2
    kn_blocks, l, N, nbconv, nblamin = 1, 0, 0, 0, 0
    lam = np.zeros(0, dtype=np.complex128)
    res = np.zeros(0)
    conv = np.zeros(0, dtype=bool)
    nlam_cols = 0
    launches0 = lib.nepb_launch_count()

    def work_blocks(m):
        if work["m"] < m:
            for key in ("wc", "bw", "z"):
                if key in work:
                    work[key].close()
            mm = max(m, min(2 * work["m"], work["m"] + 32))
            work.update(m=mm, wc=Block(n, mm), bw=Block(n, mm), z=Block(n, mm))
        return work["wc"], work["bw"], work["z"]

    def check_convergence(all_):
        nonlocal lam, res, conv, nbconv, nblamin, nlam_cols
        lambda_, S = sla.eig(K[:l, :l], H[:l, :l])
        if not all_:
            lamin = rk.in_sigma(lambda_, Sigma, tol)
            ilam = np.nonzero(lamin)[0]
            lam = lambda_[ilam]
        else:
            ilam = np.nonzero(np.isfinite(lambda_))[0]
            lam = lambda_[ilam]
            lamin = rk.in_sigma(lam, Sigma, tol)
        nblamin = int(lamin.sum())
        nlam_cols = len(ilam)
        if nlam_cols:
            HS = H[:l + 1, :l] @ S[:, ilam]
            HS = HS / np.linalg.norm(HS, axis=0)[None, :]
            block_gemm(basis.b, 0, l + 1, HS, Qb, 0, rows=n)  # X = V[1:n, 1:l+1] * (H * S[:, ilam])
            if host_err:
                X = Qb.download(0, nlam_cols)
                X = X / np.linalg.norm(X, axis=0)[None, :]
                res = np.array([errmeasure(lam[i], X[:, i]) for i in range(nlam_cols)], dtype=float)
            else:
                res = np.asarray(residual_errors(nep, errmeasure, lam, Qb, nlam_cols, Rb), dtype=float)
        else:
            res = np.zeros(0)
        conv = np.abs(res) < tol
        if all_:
            resall = np.full(l, np.nan)
            resall[ilam] = res
            order = sorted(range(l), key=lambda i: (abs(lambda_[i]), np.angle(lambda_[i])))
            Res[:l, l - 1] = resall[order]
            Lam[:l, l - 1] = lambda_[order]
            conv = conv & lamin
        nbconv = int(conv.sum()) if len(conv) else 0

    k = 1
    while k <= kmax:
        if expand:
            kn_blocks += 1
            N += 1
            nrmD.append(float(np.abs(sgdd[:, k]).max()))
            if not np.isfinite(nrmD[k]):
                raise ValueError("The generalized divided differences must be finite.")
            if n > 1 and 5 <= k < kconv:
                frozen = False
                if sum(nrmD[k - 4:k + 1]) < 5 * tollin:
                    kconv = k - 1
                    if static:
                        kmax = maxit + kconv
                        kn_blocks -= 1
                    xi, beta, nrmD = xi[:k], beta[:k], nrmD[:k]
                    frozen = True
                elif k == maxdgr + 1:
                    kconv = k
                    warnings.warn("NLEIGS: Linearization not converged after %d iterations" % maxdgr)
                    frozen = True
                if frozen:
                    expand = False
                    N -= 1
                    if leja == 1:
                        if len(sigma) < kmax + 1:
                            sigma = np.concatenate([sigma, np.zeros(kmax + 1 - len(sigma), dtype=np.complex128)])
                        sigma[k:kmax + 1] = nodes[:kmax - k + 1]
        l = k - N if static else k
        if not static or not expand:
            basis.ensure(kn_blocks, l)
            wcb, bwb, zb = work_blocks(N + 1)
            add_to_cache = ((not expand or k > kconv) and reusefact == 1) or reusefact == 2
            _device_backslash(nep, cache, basis, l, wcb, bwb, zb, tb, sigma, k, beta, N, xi, sgdd, add_to_cache)
            h, nrm, _ = dgks(basis.b, l, basis.b, l, rows=n * kn_blocks)
            H[:l, l - 1] = h
            H[l, l - 1] = nrm
            K[:l, l - 1] = h * sigma[k]
            K[l - 1, l - 1] += 1.0
            K[l, l - 1] = nrm * sigma[k]
        if not return_details and ((not expand and k >= N + minit and (k - (N + minit)) % check_error_every == 0) or
                                   (k >= kconv + minit and (k - (kconv + minit)) % check_error_every == 0) or k == kmax):
            check_convergence(False)
        elif return_details and (not static or not expand):
            check_convergence(True)
        if ((not expand and k >= N + minit) or k >= kconv + minit) and nblamin == nbconv:
            break
        k += 1

    X = Qb.download(0, nlam_cols) if nlam_cols else np.zeros((n, 0), dtype=np.complex128)
    if nlam_cols:
        X = X / np.linalg.norm(X, axis=0)[None, :]
    details = {"Lam": Lam[:l, :l], "Res": Res[:l, :l], "sigma": sigma[:min(k, len(sigma))], "xi": xi[:k] if expand else xi,
               "beta": beta[:k] if expand else beta, "nrmD": nrmD[:k] if expand else nrmD, "kconv": kconv, "iterations": min(k, kmax),
               "factorizations": len(cache.solvers), "N": N, "l": l, "H": H[:l + 1, :l], "K": K[:l + 1, :l],
               "gpu_launches": int(lib.nepb_launch_count() - launches0), "lam_all": lam, "res_all": res}
    if return_details and expand:
        warnings.warn("NLEIGS: Linearization not converged after %d iterations" % maxdgr)
    for blk in [basis, Qb, Rb, tb] + [work[key] for key in ("wc", "bw", "z") if key in work]:
        blk.close()
    for s in cache.solvers.values():
        _release(s)
    return lam[conv], X[:, conv], res[conv], details


# ---------------------------------------------------------------------------------------------
# low-rank branch: SumNEP(PEP, LowRankFactorizedNEP) as in the reference's gun variants R1 / R2 / S
# ---------------------------------------------------------------------------------------------
class LowRankStructure:
    """What get_rk_nep attaches for `SumNEP(PEP, LowRankFactorizedNEP)` (rk_helper/rk_nep.jl:127-152): polynomial degree p, the
    L factors of the q nonlinear terms side by side (n x r), UU = hcat(U...) and the term every column belongs to."""

    def __init__(self, p, L, U):
        import scipy.sparse as sp
        self.p, self.q = int(p), len(L)
        self.L = [sp.csr_matrix(x) for x in L]
        self.Lcat = sp.hstack(self.L).tocsr()
        self.UUt = sp.hstack([sp.csr_matrix(u) for u in U]).conj().T.tocsr()  # r x n
        self.r = self.Lcat.shape[1]
        self.iL = np.concatenate([np.full(x.shape[1], i) for i, x in enumerate(self.L)])

    @classmethod
    def from_terms(cls, p, nonlinear_matrices):
        LU = [rk.low_rank_lu_factors(A) for A in nonlinear_matrices]
        return cls(p, [x[0] for x in LU], [x[1] for x in LU])


def lowrank_backslash(nep: B200SPMF, P: LowRankStructure, solve, wc, sigma, k, beta, N, xi, sgdd):
    """`backslash` of method_nleigs.jl:399-518 with `P.is_low_rank`: after the p-th block the blocks of the continuation vector
    have r entries (sum of the ranks) instead of n, so everything but the first p blocks is tiny and stays on the host; the
    device does what is O(nnz) or worse -- the products with all A_i at once (`P.BBCC * block` weighted by a column of sgdd =
    ONE fused SpMM in SCALAR mode with C_i = sgdd[i, .]) and the shifted solve.  0-based indices as in the oracle."""
    n, p, r = nep.n, P.p, P.r
    shift = sigma[k]
    wc = np.asarray(wc, dtype=np.complex128)

    def stacked(block, col):  # sum_i sgdd[i, col] A_i block
        return nep.apply(_lib.COEF_SCALAR, block.reshape(n, 1), np.ascontiguousarray(sgdd[:, col]), 1)[:, 0]

    Bw = np.zeros_like(wc)
    Bw[:n] = -stacked(wc[(p - 1) * n:p * n], p) / beta[p]  # first block (:408-416)
    i0b, i0e = 0, n
    for ii in range(1, N + 1):  # other blocks (:418-435)
        i1b, i1e = i0e, i0e + (n if ii < p else r)
        if ii != p:
            Bw[i1b:i1e] = wc[i0b:i0e] + beta[ii] / xi[ii - 1] * wc[i1b:i1e]
        else:
            Bw[i1b:i1e] = P.UUt @ wc[i0b:i0e] + beta[ii] / xi[ii - 1] * wc[i1b:i1e]
        i0b, i0e = i1b, i1e
    z = Bw.copy()  # z0 (:437-489)
    i1b, i1e = n, (2 * n if p > 1 else n + r)
    z[i1b:i1e] = z[i1b:i1e] / (beta[1] * (1 - shift / xi[0]))
    for ii in range(1, N + 1):
        i2b, i2e = i1e, i1e + (n if ii < p - 1 else r)
        if ii < p:
            z[:n] -= stacked(z[i1b:i1e], ii)
        elif ii > p:
            z[:n] -= P.Lcat @ (z[i1b:i1e] * sgdd[p + 1:, ii][P.iL])  # the LL / iLr loops (:463-470)
        if ii < N:
            mu = shift - sigma[ii]
            nu = beta[ii + 1] * (1 - shift / xi[ii])
            if ii != p - 1:
                z[i2b:i2e] = z[i2b:i2e] / nu + mu / nu * z[i1b:i1e]
            else:
                z[i2b:i2e] = z[i2b:i2e] / nu + mu / nu * (P.UUt @ z[i1b:i1e])
        i1b, i1e = i2b, i2e
    w = np.zeros_like(wc)  # solve and substitutions (:491-515)
    w[:n] = solve(shift, z[:n] / beta[0])
    i0b, i0e = 0, n
    for ii in range(1, N + 1):
        i1b, i1e = i0e, i0e + (n if ii < p else r)
        mu = shift - sigma[ii - 1]
        nu = beta[ii] * (1 - shift / xi[ii - 1])
        if ii != p:
            w[i1b:i1e] = mu / nu * w[i0b:i0e] + Bw[i1b:i1e] / nu
        else:
            w[i1b:i1e] = mu / nu * (P.UUt @ w[i0b:i0e]) + Bw[i1b:i1e] / nu
        i0b, i0e = i1b, i1e
    return w


def nleigs_lowrank(nep: B200SPMF, lowrank: LowRankStructure, Sigma=(-1.0 - 1j, -1 + 1j, 1 + 1j, 1 - 1j), Xi=(np.inf,), maxdgr=100,
                   minit=20, maxit=200, tol=1e-10, tollin=None, v=None, errmeasure=None, isfunm=True, static=False, leja=1, nodes=(),
                   reusefact=1, check_error_every=5, umfpack_refinements=0, linsolvercache=None):
    """nleigs (method_nleigs.jl:60-377) for `SumNEP(PEP, LowRankFactorizedNEP)`: `nep` is the device operator with the terms in
    the order (PEP coefficients 0..p, nonlinear terms), `lowrank` their low-rank structure.  The rational Krylov vectors have
    p n + (N - p + 1) r entries (gun: 9956 + 84 per extra block instead of 9956 per block), so the basis, DGKS and the small
    recurrences live on the host; the device does the products with all A_i and the cached shifted solves.
    Returns (lam, X, res, details)."""
    from .solvers import dgks_host
    import scipy.linalg as sla
    P = lowrank
    Sigma = np.asarray(Sigma, dtype=np.complex128)
    Xi = np.asarray(Xi, dtype=np.float64)
    n, p_poly, r = nep.n, P.p, P.r
    tollin = max(tol / 10, 100 * np.finfo(float).eps) if tollin is None else tollin
    v = np.random.default_rng(0).standard_normal(n) if v is None else v
    v = np.asarray(v, dtype=np.complex128)
    errmeasure = errmeasure or ResidualErrmeasure(nep)
    host_err = callable(errmeasure) and not hasattr(errmeasure, "estimate_error")
    nodes = np.asarray(nodes, dtype=np.complex128)
    if leja == 0:  # nodes, poles, scalings (:120-146)
        if len(nodes) == 0:
            raise ValueError("Interpolation nodes must be provided via 'nodes' when no Leja-Bagby points ('leja' == 0) are used.")
        gamma, _ = rk.discretizepolygon(Sigma)
        max_count = maxit + maxdgr + 2 if static else max(maxit, maxdgr) + 2
        sigma = np.tile(nodes, -(-max_count // len(nodes)))
        _, xi, beta = rk.lejabagby(sigma[:maxdgr + 2], Xi, gamma, maxdgr + 2, True, p_poly)
    elif leja == 1:
        if len(nodes) == 0:
            gamma, nodes = rk.discretizepolygon(Sigma, True)
        else:
            gamma, _ = rk.discretizepolygon(Sigma)
        nodes = np.tile(nodes, -(-(maxit + 1) // len(nodes)))
        sigma, xi, beta = rk.lejabagby(gamma, Xi, gamma, maxdgr + 2, False, p_poly)
    else:
        gamma, _ = rk.discretizepolygon(Sigma)
        max_count = maxit + maxdgr + 2 if static else max(maxit, maxdgr) + 2
        sigma, xi, beta = rk.lejabagby(gamma, Xi, gamma, max_count, False, p_poly)
    sigma = np.array(sigma, dtype=np.complex128)
    xi = np.array(xi, dtype=np.float64)
    xi[maxdgr + 1] = np.nan
    head = slice(0, maxdgr + 2)
    sgdd = rk.scgendivdiffs(sigma[head], xi[head], beta[head], maxdgr, isfunm, nep.get_fv())
    nrmD = [float(np.abs(sgdd[:, 0]).max())]
    if not np.isfinite(nrmD[0]):
        raise ValueError("The generalized divided differences must be finite.")
    kmax = maxit + maxdgr if static else maxit
    cache = linsolvercache or DeviceLinSolverCache(nep, umfpack_refinements)  # (`linsolvercache`: tests inject a host stand-in)

    def solve(shift, y, add_to_cache):
        s = cache.get(shift, add_to_cache)
        x = s.lin_solve(y)
        if cache.solvers.get(complex(shift)) is not s:
            _release(s)
        return x

    launches0 = lib.nepb_launch_count()
    v = solve(sigma[0], v / np.linalg.norm(v), reusefact == 2)
    V = np.zeros((p_poly * n + (kmax + 2) * r + n, kmax + 1), dtype=np.complex128)
    V[:n, 0] = v / np.linalg.norm(v)
    H = np.zeros((kmax + 1, kmax), dtype=np.complex128)
    K = np.zeros((kmax + 1, kmax), dtype=np.complex128)
    expand, kconv = True, np.iinfo(np.int64).max // 2
    kn, l, N, nbconv, nblamin = n, 0, 0, 0, 0
    lam = np.zeros(0, dtype=np.complex128)
    X = np.zeros((n, 0), dtype=np.complex128)
    res = np.zeros(0)
    conv = np.zeros(0, dtype=bool)
    nfact = 0
    k = 1
    while k <= kmax:
        if expand:
            kn += n if k < p_poly else r  # (:205-211)
            N += 1
            nrmD.append(float(np.abs(sgdd[:, k]).max()))
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
                    xi, beta, nrmD = xi[:k], beta[:k], nrmD[:k]
                    if static:
                        kn -= n if k < p_poly else r
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
            if kn > V.shape[0] or l + 1 > V.shape[1]:
                W = np.zeros((max(kn, V.shape[0]), max(l + 1, V.shape[1])), dtype=np.complex128)
                W[:V.shape[0], :V.shape[1]] = V
                V = W
            t = np.zeros(l, dtype=np.complex128)
            t[l - 1] = 1
            add_to_cache = ((not expand or k > kconv) and reusefact == 1) or reusefact == 2
            w = lowrank_backslash(nep, P, lambda s, y: solve(s, y, add_to_cache), V[:kn, l - 1].copy(), sigma, k, beta, N, xi, sgdd)
            h = np.zeros(l, dtype=np.complex128)
            H[l, l - 1] = dgks_host(V[:kn, :l], w, h)
            H[:l, l - 1] = h
            K[:l, l - 1] = h * sigma[k] + t
            K[l, l - 1] = H[l, l - 1] * sigma[k]
            V[:kn, l] = w
        check = ((not expand and k >= N + minit and (k - (N + minit)) % check_error_every == 0) or
                 (k >= kconv + minit and (k - (kconv + minit)) % check_error_every == 0) or k == kmax)
        if check:
            lambda_, S = sla.eig(K[:l, :l], H[:l, :l])
            lamin = rk.in_sigma(lambda_, Sigma, tol)
            ilam = np.nonzero(lamin)[0]
            lam = lambda_[ilam]
            nblamin = int(np.sum(lamin))
            for i in ilam:
                S[:, i] = S[:, i] / np.linalg.norm(H[:l + 1, :l] @ S[:, i])
            X = V[:n, :l + 1] @ (H[:l + 1, :l] @ S[:, ilam])
            X = X / np.linalg.norm(X, axis=0)[None, :] if X.shape[1] else X
            if len(lam) == 0:
                res = np.zeros(0)
            elif host_err:
                res = np.array([errmeasure(lam[i], X[:, i]) for i in range(len(lam))], dtype=float)
            else:
                res = np.asarray(errmeasure.estimate_errors(lam, X) if hasattr(errmeasure, "estimate_errors")
                                 else [errmeasure.estimate_error(lam[i], X[:, i]) for i in range(len(lam))], dtype=float)
            conv = np.abs(res) < tol
            nbconv = int(np.sum(conv)) if len(conv) else 0
        if ((not expand and k >= N + minit) or k >= kconv + minit) and nblamin == nbconv:
            break
        k += 1
    details = {"sigma": sigma[:min(k, len(sigma))], "xi": xi, "beta": beta, "nrmD": nrmD, "kconv": kconv, "iterations": min(k, kmax),
               "factorizations": len(cache.solvers), "N": N, "l": l, "rows": kn, "gpu_launches": int(lib.nepb_launch_count() - launches0),
               "lam_all": lam, "res_all": res}
    for s in cache.solvers.values():
        _release(s)
    return lam[conv], X[:, conv], res[conv], details
