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
from .neptypes import B200SPMF, Block
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
    solve_block(solver.lu, t, 0, 1, bwb, 0, alpha=1.0 / beta[0])  # Bw_0 is unused from here on: it receives w0
    block_gemm(bwb, 0, m, CW, wcb, 0)
    w = wcb.download()
    for b in (wcb, bwb, zb, t):
        b.close()
    return w.reshape(n * m, order="F")
