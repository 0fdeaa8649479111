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
