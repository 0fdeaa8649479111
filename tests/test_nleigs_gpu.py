"""nleigs' `backslash` (src/method_nleigs.jl:399-518, full-rank SPMF branch) on the device against the oracle restatement,
on the gun problem (n = 9956 > 400, so the reference takes the stacked-BBCC path) and on the synthetic degree-3 PEP."""
import numpy as np
import pytest
import scipy.sparse as sp

import nepb200
from nepb200 import B200SPMF, ONE, IDENTITY, PowShift, Monomial
from oracle import gallery as g
from oracle import nep as o
from oracle import nleigs as onl
from oracle import solvers as osol

pytestmark = pytest.mark.gpu


def _scalars(N, p, rng, centre, radius):
    sigma = centre + radius * np.exp(2j * np.pi * rng.random(N + 2))      # interpolation nodes / shifts
    xi = centre + 3.0 * radius * np.exp(2j * np.pi * rng.random(N + 2))   # poles outside the target set
    beta = 0.5 + rng.random(N + 2)
    sgdd = (rng.standard_normal((p, N + 2)) + 1j * rng.standard_normal((p, N + 2))) / (1.0 + np.arange(N + 2))[None, :]
    return sigma, xi, beta, sgdd


@pytest.mark.parametrize("N,k", [(1, 0), (3, 2), (8, 5)])
def test_backslash_gun(N, k):
    K, M, W1, W2 = g.load_gun_matrices()
    Av = [K, -M, W1, W2]
    dnep = B200SPMF(Av, [ONE, IDENTITY, PowShift(0.5, 0.0, 1j), PowShift(0.5, 108.8774 ** 2, 1j)])
    onep = o.nep_gallery("nlevp_native_gun")
    n = dnep.n
    rng = np.random.default_rng(N)
    sigma, xi, beta, sgdd = _scalars(N, 4, rng, 250.0 ** 2, 5e4)
    wc = rng.standard_normal(n * (N + 1)) + 1j * rng.standard_normal(n * (N + 1))
    solvers = {}

    def solve(shift, rhs):
        if shift not in solvers:
            solvers[shift] = osol.FactorizeLinSolver(onep, shift)
        return solvers[shift].lin_solve(rhs)

    wo = onl.backslash_fullrank(wc, Av, solve, sigma, k, beta, N, xi, sgdd)
    cache = nepb200.DeviceLinSolverCache(dnep)
    w = nepb200.nleigs_backslash(dnep, cache, wc, sigma, k, beta, N, xi, sgdd)
    assert np.linalg.norm(w - wo) <= 1e-10 * np.linalg.norm(wo)
    assert len(cache.solvers) == 1
    nepb200.nleigs_backslash(dnep, cache, wc, sigma, k, beta, N, xi, sgdd, add_to_cache=False)
    assert len(cache.solvers) == 1  # the cached factorisation is reused (method_nleigs.jl:490-491)


def test_backslash_stencil_pep():
    from nepb200 import synthetic
    mats, _ = synthetic.stencil_pep(40)
    Av = [m.tocsc() for m in mats]
    dnep = B200SPMF(Av, [Monomial(i) for i in range(4)])
    n = dnep.n
    N, k = 4, 3
    rng = np.random.default_rng(7)
    sigma, xi, beta, sgdd = _scalars(N, 4, rng, 0.0, 0.5)
    wc = rng.standard_normal(n * (N + 1)) + 0j

    def solve(shift, rhs):
        Mo = sum(A * shift ** i for i, A in enumerate(Av)).tocsc()
        import scipy.sparse.linalg as sla
        return sla.splu(sp.csc_matrix(Mo, dtype=complex)).solve(rhs.astype(complex))

    wo = onl.backslash_fullrank(wc, Av, solve, sigma, k, beta, N, xi, sgdd)
    w = nepb200.nleigs_backslash(dnep, nepb200.DeviceLinSolverCache(dnep), wc, sigma, k, beta, N, xi, sgdd)
    assert np.linalg.norm(w - wo) <= 1e-10 * np.linalg.norm(wo)
