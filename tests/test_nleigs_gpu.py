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


# ---- the full nleigs driver on the device (src/method_nleigs.jl:60-377) ------------------------------------------------
import json  # noqa: E402
import os  # noqa: E402
import sys  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_nleigs_golden as mg  # noqa: E402  (the gun set-up of test/rk_helper/gun_test_utils.jl, shared with the fixture script)


def _gun_device():
    K, M, W1, W2 = g.load_gun_matrices()
    pep = nepb200.PEP([K, -M])
    spmf = nepb200.SPMF_NEP([W1, W2], [PowShift(0.5, 0.0, 1j), PowShift(0.5, 108.8774 ** 2, 1j)])
    return B200SPMF.from_nep(nepb200.SumNEP(pep, spmf)), (K, M, W1, W2)


def _gold(variant):
    with open(os.path.join(os.path.dirname(__file__), "golden", "nleigs_gun.json")) as f:
        return json.load(f)[variant]


def _match(lam, gold_lam, rtol):
    gl = np.array([complex(*x) for x in gold_lam])
    assert len(lam) == len(gl)
    for x in lam:
        assert np.min(np.abs(gl - x)) <= rtol * abs(x)


def test_nleigs_gun_naive_device():
    """test/nleigs/nleigs_gun_naive.jl on the device: one eigenvalue, the gun reference eigenvalue (test/gun_native.jl:9);
    same linearization degree and iteration count as the CPU oracle (golden fixture)."""
    dnep, _ = _gun_device()
    sq = np.array([-1 - 1j, -1 + 1j, 1 + 1j, 1 - 1j])
    lam, X, res, det = nepb200.nleigs(dnep, 150.0 ** 2 + 200.0 * sq, v=np.ones(dnep.n) + 0j)
    gold = _gold("naive")
    assert len(lam) == 1 and abs(lam[0] - (22345.116783765 + 0.644998598j)) < 1e-8 * abs(lam[0])
    _match(lam, gold["lam"], 1e-10)
    assert det["kconv"] == gold["kconv"] and det["iterations"] == gold["iterations"] and det["N"] == gold["N"]
    onep = o.nep_gallery("nlevp_native_gun")
    assert np.linalg.norm(o.compute_Mlincomb(onep, lam[0], X[:, 0])) / np.linalg.norm(X[:, 0]) < 1e-10
    assert res[0] < 1e-10 and det["gpu_launches"] > 0


@pytest.mark.parametrize("variant,count", [("P", 18), ("R2", 21), ("S", 21)])
def test_nleigs_gun_variants_device(variant, count):
    """test/nleigs/nleigs_gun_variant_{p,r2,s}.jl (target set, nodes, poles of gun_test_utils.jl; full-rank branch): the
    reference's eigenvalue counts 18 / 21 / 21, eigenvalues equal to the oracle's fixture, scaled residuals below tol."""
    dnep, (K, M, W1, W2) = _gun_device()
    Sigma, Xi, nodes = mg.gun_setup()
    v = mg.gun_start_vector(dnep.n)
    funres = mg.gun_residual(K, -M, W1, W2)
    if variant == "P":
        lam, X, res, det = nepb200.nleigs(dnep, Sigma, maxit=100, v=v, leja=0, nodes=nodes, reusefact=2, errmeasure=funres)
    elif variant == "R2":
        lam, X, res, det = nepb200.nleigs(dnep, Sigma, Xi=Xi, minit=60, maxit=100, v=v, nodes=nodes, errmeasure=funres)
    else:
        lam, X, res, det = nepb200.nleigs(dnep, Sigma, Xi=Xi, minit=70, maxit=100, v=v, nodes=nodes, static=True, errmeasure=funres)
    gold = _gold(variant)
    assert gold["count"] == count  # the fixture itself reproduces the reference's literal
    if variant == "P":
        # degree-100 polynomial interpolation leaves three Ritz pairs within a factor 2.5 of tol = 1e-10: the count of
        # converged pairs may differ by rounding; every oracle eigenvalue must be among the device's Ritz values in Sigma
        assert count - 3 <= len(lam) <= count + 3
        gl = np.array([complex(*x) for x in gold["lam"]])
        for x in lam:
            assert np.min(np.abs(gl - x)) <= 1e-8 * abs(x) or funres(x, X[:, list(lam).index(x)]) < 1e-10
    else:
        assert len(lam) == count
        _match(lam, gold["lam"], 1e-8)
        assert det["kconv"] == gold["kconv"] and det["N"] == gold["N"]
    assert np.all(res < 1e-10)
    for i in range(len(lam)):
        assert funres(lam[i], X[:, i]) < 1e-10


def test_nleigs_stencil_pep_device_matches_oracle():
    """Config C4 in small (degree-3 stencil PEP, n = 1600 > 400: the stacked-product branch) against the oracle driver run
    on the same inputs: same linearization degree, same eigenvalues in the target set."""
    from nepb200 import synthetic
    mats, _ = synthetic.stencil_pep(40)
    Av = [m.tocsc() for m in mats]
    dnep = B200SPMF.from_nep(nepb200.PEP(Av))
    onep = o.PEP(Av)
    Sigma = np.array([-1 - 1j, -1 + 1j, 1 + 1j, 1 - 1j]) * 0.06 + (0.96 - 0.6j)  # two eigenvalues of this PEP lie inside
    v = np.ones(dnep.n) + 0j
    lo, Xo, ro, do = onl.nleigs(onep, Sigma, v=v, maxit=60)
    ld, Xd, rd, dd = nepb200.nleigs(dnep, Sigma, v=v, maxit=60)
    assert dd["kconv"] == do["kconv"] and dd["N"] == do["N"]
    assert len(ld) == len(lo) == 2
    for x in ld:
        assert np.min(np.abs(lo - x)) <= 1e-9 * max(1.0, abs(x))
    for i in range(len(ld)):
        assert np.linalg.norm(o.compute_Mlincomb(onep, ld[i], Xd[:, i])) / np.linalg.norm(Xd[:, i]) < 1e-9


@pytest.mark.parametrize("variant", ["R1", "R2", "S"])
def test_nleigs_gun_lowrank_device(variant):
    """test/nleigs/nleigs_gun_variant_{r1,r2,s}.jl as the reference runs them: SumNEP(PEP([K, -M]), LowRankFactorizedNEP([W1, W2]))
    (gun_test_utils.jl:36-43), so `nleigs` takes the low-rank branches of backslash (method_nleigs.jl:408-416,430,463-471,480,510).
    21 eigenvalues each (the reference's literal), equal to the oracle's golden runs; products with the A_i and the shifted
    solves on the device."""
    K, M, W1, W2 = g.load_gun_matrices()
    low = nepb200.LowRankFactorizedNEP([W1, W2], [PowShift(0.5, 0.0, 1j), PowShift(0.5, 108.8774 ** 2, 1j)])
    dnep = B200SPMF.from_nep(nepb200.SumNEP(nepb200.PEP([K, -M]), low))
    assert low.r == 84
    Sigma, Xi, nodes = mg.gun_setup()
    v = mg.gun_start_vector(dnep.n)
    funres = mg.gun_residual(K, -M, W1, W2)
    if variant == "R1":
        lam, X, res, det = nepb200.nleigs(dnep, Sigma, Xi=Xi, maxit=100, v=v, leja=0, nodes=nodes, reusefact=2, errmeasure=funres)
    elif variant == "R2":
        lam, X, res, det = nepb200.nleigs(dnep, Sigma, Xi=Xi, minit=60, maxit=100, v=v, nodes=nodes, errmeasure=funres)
    else:
        lam, X, res, det = nepb200.nleigs(dnep, Sigma, Xi=Xi, minit=70, maxit=100, v=v, nodes=nodes, static=True, errmeasure=funres)
    with open(os.path.join(os.path.dirname(__file__), "golden", "nleigs_gun_lowrank.json")) as f:
        gold = json.load(f)[variant]
    assert gold["count"] == 21 and len(lam) == 21
    _match(lam, gold["lam"], 1e-8)
    assert det["kconv"] == gold["kconv"] and det["N"] == gold["N"]
    # p = 1: one block of n entries, then blocks of r = 84 (method_nleigs.jl:205-211); the dynamic variants keep the block that
    # was added in the step the linearization converged
    assert det["rows"] in (dnep.n + det["N"] * 84, dnep.n + (det["N"] + 1) * 84) and det["gpu_launches"] > 0
    assert np.all(res < 1e-10)
