"""nleigs: the oracle restatement (oracle/nleigs.py) pinned to the literals of the reference's own tests, and the product's
host helpers (nepb200.rk_helper) checked against the same literals and against the oracle, bit for bit.

Reference tests followed: test/rk_helper/inpolygon.jl, test/rk_helper/discretizepolygon.jl, test/nleigs/nleigs_basic.jl,
test/nleigs/nleigs_scalar.jl, test/nleigs/nleigs_gun_naive.jl, the nleigs docstring (src/method_nleigs.jl:42-51) and the PEP
example of docs/src/index.md:40-42."""
import warnings

import numpy as np
import pytest
import scipy.linalg as sla

import nepb200
from nepb200 import rk_helper as prk
from oracle import nep as o
from oracle import nleigs as onl

POLYX = [0, 0, 5, 10, 10]
POLYY = [0, 10, 5, 10, 0]


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_inpolygon_reference_counts(impl):
    pts = np.array([complex(x, y) for x in range(-1, 12) for y in range(-1, 12)])
    for px, py in ((POLYX, POLYY), (POLYX[::-1], POLYY[::-1])):
        if impl == "oracle":
            inside = [onl.inpolygon(p.real, p.imag, px, py) for p in pts]
        else:
            inside = prk.inpolygon(pts.real, pts.imag, px, py)
        assert len(inside) == 13 * 13 and int(np.sum(inside)) == 96  # test/rk_helper/inpolygon.jl:13-17
    for bad in (complex(np.nan, 0), complex(0, np.nan), complex(np.inf, 0), complex(0, np.inf)):
        if impl == "oracle":
            assert not onl.inpolygon(bad.real, bad.imag, POLYX, POLYY)
        else:
            assert not prk.inpolygon(bad.real, bad.imag, POLYX, POLYY)[0]


EXPECTED_BOUNDARY = np.array([
    0 + 0j, 0 + 2.2071068j, 0 + 4.4142136j, 0 + 6.6213203j, 0 + 8.8284271j, 0.73223305 + 9.267767j, 2.2928932 + 7.7071068j,
    3.8535534 + 6.1464466j, 5.4142136 + 5.4142136j, 6.9748737 + 6.9748737j, 8.5355339 + 8.5355339j, 10 + 9.863961j, 10 + 7.6568542j,
    10 + 5.4497475j, 10 + 3.2426407j, 10 + 1.0355339j, 8.8284271 + 0j, 6.6213203 + 0j, 4.4142136 + 0j, 2.2071068 + 0j])


@pytest.mark.parametrize("mod", [onl, prk], ids=["oracle", "product"])
def test_discretizepolygon_reference_cases(mod):
    poly = np.array([0.0, 10j, 5 + 5j, 10 + 10j, 10 + 0j])
    boundary, interior = mod.discretizepolygon(poly, True, 20, 100)
    assert len(boundary) == 20 + len(poly) + 1
    assert np.allclose(boundary[:20], EXPECTED_BOUNDARY, rtol=1e-7, atol=1e-7)  # discretizepolygon.jl:8-24 (8 printed digits)
    assert np.array_equal(boundary[20:25], poly) and boundary[25] == poly[0]
    assert len(interior) >= 100
    assert all(onl.inpolygon(p.real, p.imag, poly.real, poly.imag) for p in interior)
    b, i = mod.discretizepolygon([-10.0 - 2j, 10 - 2j, 10 + 2j, -10 + 2j], True, 100, 5)  # narrow
    assert len(b) == 105 and len(i) >= 5
    with pytest.raises(RuntimeError):
        mod.discretizepolygon([-10.0 - 0.2j, 10 - 0.2j, 10 + 0.2j, -10 + 0.2j], True, 100, 5)  # too narrow
    b, i = mod.discretizepolygon([], True, 100, 100)  # unit disk
    assert len(b) == 101 and np.allclose(np.abs(b[:100]), 1.0, rtol=100 * np.finfo(float).eps)
    assert len(i) >= 100 and np.all(np.abs(i) < 1)
    p1, p2 = -2.0 - 1j, 2.0 + 1j
    b, i = mod.discretizepolygon([p1, p2], True, 100, 100)  # Chebyshev points
    assert len(b) == 102 and np.all(np.abs(((b - p1) / (p2 - p1)).imag) < 1e-15)
    assert len(i) >= 100 and np.all(np.abs(((i - p1) / (p2 - p1)).imag) < 1e-15)


def test_product_helpers_equal_oracle_bitwise():
    """Integer / selection work (greedy Leja-Bagby choices, polygon membership) must agree exactly; the floating-point
    sequences are produced by the same operations in the same order, so they agree bit for bit as well."""
    Sigma = np.array([-1 - 1j, -1 + 1j, 1 + 1j, 1 - 1j]) * 200.0 + 150.0 ** 2
    g1, n1 = onl.discretizepolygon(Sigma, True)
    g2, n2 = prk.discretizepolygon(Sigma, True)
    assert np.array_equal(g1, g2) and np.array_equal(n1, n2)
    Xi = -10 ** np.linspace(-8, 8, 10000) + 108.8774 ** 2
    for xi_set, force in ((np.array([np.inf]), 1), (Xi, 1), (Xi, 0), (Xi, 3)):
        a1, b1, c1 = onl.lejabagby(g1, xi_set, g1, 40, False, force)
        a2, b2, c2 = prk.lejabagby(g1, xi_set, g1, 40, False, force)
        assert np.array_equal(a1, a2) and np.array_equal(b1, b2) and np.array_equal(c1, c2)
    nodes = np.tile(150.0 ** 2 + 100.0 * np.array([2 / 3, (1 + 1j) / 3, 0, (-1 + 1j) / 3, -2 / 3]), 8)
    a1, b1, c1 = onl.lejabagby(nodes, Xi, g1, 40, True, 1)
    a2, b2, c2 = prk.lejabagby(nodes, Xi, g1, 40, True, 1)
    assert np.array_equal(a1, a2) and np.array_equal(b1, b2) and np.array_equal(c1, c2)
    z = g1[::97] * (1 + 1e-3 * np.exp(1j * np.arange(len(g1[::97]))))
    assert np.array_equal(onl.in_sigma(z, Sigma, 1e-10), prk.in_sigma(z, Sigma, 1e-10))
    assert np.array_equal(onl.in_sigma([0.5, 4.1, 2 + 1e-12j], [0.01 + 0j, 4.0], 1e-10), prk.in_sigma([0.5, 4.1, 2 + 1e-12j], [0.01 + 0j, 4.0], 1e-10))
    # scalar generalized divided differences: oracle callables vs the product's function classes
    sg, xg, bg = a1[:12], b1[:12], c1[:12]
    ofv = [o.f_one, o.f_id, o.f_isqrt_shift(0.0), o.f_isqrt_shift(108.8774 ** 2)]
    pfv = [nepb200.ONE, nepb200.IDENTITY, nepb200.PowShift(0.5, 0.0, 1j), nepb200.PowShift(0.5, 108.8774 ** 2, 1j)]
    d1 = onl.scgendivdiffs(sg, xg, bg, 10, True, ofv)
    d2 = prk.scgendivdiffs(sg, xg, bg, 10, True, pfv)
    assert np.allclose(d1, d2, rtol=1e-13, atol=1e-300)
    ag, bgr, cgr = onl.lejabagby(g1, Xi, g1, 8, False, 1)  # greedy nodes are distinct: differencing applies
    su = ag
    assert len(np.unique(su)) == 8
    d3 = onl.scgendivdiffs(su, bgr, cgr, 6, False, ofv)
    d4 = prk.scgendivdiffs(su, bgr, cgr, 6, False, pfv)
    # differencing cancels: trailing coefficients carry absolute errors of a few ulps of the leading ones
    assert np.allclose(d3, d4, rtol=1e-9, atol=1e-11 * np.abs(d3).max())
    d5 = prk.scgendivdiffs(su, bgr, cgr, 6, True, pfv)  # matrix-function route == differencing route
    assert np.allclose(d4, d5, rtol=1e-7, atol=1e-12 * np.abs(d4).max())


def test_rk_structure():
    import scipy.sparse as sp
    A = [sp.identity(4, format="csc")] * 3
    pep = nepb200.PEP(A)
    spmf = nepb200.SPMF_NEP(A[:2], [nepb200.PowShift(0.5, 0.0, 1j), nepb200.Exp(-1.0)])
    assert prk.rk_structure(pep, pep) == (2, 0)
    s = nepb200.SumNEP(pep, spmf)
    assert prk.rk_structure(s, s) == (2, 2)
    assert prk.rk_structure(spmf, spmf) == (-1, 2)
    assert onl.RKNEP(o.PEP([np.eye(2)] * 3)).p == 2
    gun_like = o.SumNEP(o.PEP([np.eye(2)] * 2), o.SPMF_NEP([np.eye(2)] * 2, [o.f_isqrt_shift(0.0), o.f_isqrt_shift(1.0)]))
    assert (onl.RKNEP(gun_like).p, onl.RKNEP(gun_like).q) == (1, 2)


# ---- the driver: test/nleigs/nleigs_basic.jl ------------------------------------------------------------------------
B = [np.array([[1.0, 3], [5, 6]]), np.array([[3.0, 4], [6, 6]]), np.eye(2)]
SIGMA_BASIC = [-10.0 - 2j, 10 - 2j, 10 + 2j, -10 + 2j]


def _verify(nep, lam, X, count, tol=1e-10):
    assert len(lam) == count
    for i in range(len(lam)):
        assert np.linalg.norm(o.compute_Mlincomb(nep, lam[i], X[:, i])) / np.linalg.norm(X[:, i]) < tol * 100


def test_nleigs_basic_polynomial():
    pep = o.PEP(B)
    lam, X, res, det = onl.nleigs(pep, SIGMA_BASIC, maxit=10, v=np.ones(2) + 0j, blksize=5)
    _verify(pep, lam, X, 4)
    # docs/src/index.md:40-42 prints these four eigenvalues of the same PEP
    doc = np.array([1.36267, -0.824084 + 0.280682j, -0.824084 - 0.280682j, -8.7145])
    for d in doc:
        assert np.min(np.abs(lam - d)) < 1e-5


def test_nleigs_basic_nonconvergent_linearization():
    pep = o.PEP(B)
    for static in (False, True):
        with pytest.warns(UserWarning, match="Linearization not converged"):
            lam, X, _, _ = onl.nleigs(pep, SIGMA_BASIC, maxit=10, v=np.ones(2) + 0j, maxdgr=5, blksize=5, static=static)
        _verify(pep, lam, X, 4)
    with pytest.warns(UserWarning, match="Linearization not converged"):
        lam, X, _, _ = onl.nleigs(pep, SIGMA_BASIC, maxit=5, v=np.ones(2) + 0j, blksize=5, return_details=True)
    assert len(lam) == 0


def test_nleigs_basic_complex_and_details():
    cpep = o.PEP([b + 1j * np.eye(2) for b in B])
    lam, X, _, _ = onl.nleigs(cpep, SIGMA_BASIC, maxit=10, v=np.ones(2) + 0j, blksize=5, return_details=True)
    _verify(cpep, lam, X, 3)
    pep = o.PEP(B)
    lam, X, _, _ = onl.nleigs(pep, SIGMA_BASIC, maxit=10, v=np.ones(2) * (1 + 0.1j), blksize=5, return_details=True)
    _verify(pep, lam, X, 4)
    lam, X, res, det = onl.nleigs(pep, SIGMA_BASIC, maxit=10, v=np.ones(2) + 0j, blksize=5, return_details=True)
    _verify(pep, lam, X, 4)
    L, R = det["Lam"][:, -1], det["Res"][:, -1]
    conv = [R[i] < 1e-12 and onl.inpolygon(L[i].real, L[i].imag, np.real(SIGMA_BASIC), np.imag(SIGMA_BASIC)) for i in range(len(L))]
    lamconv = L[conv]
    assert len(lamconv) == 4
    assert all(np.min(np.abs(lam - x)) < 1e-12 * max(1, abs(x)) for x in lamconv)


def test_nleigs_dep0_docstring():
    """src/method_nleigs.jl:42-51: eigenpairs of dep0 in the unit square with residual norms ~1e-13; they are the three
    eigenvalues Beyn's method finds in the same region (src/method_beyncontour.jl:36-42)."""
    dep = o.nep_gallery("dep0")
    lam, X, res, _ = onl.nleigs(dep, [1 + 1j, 1 - 1j, -1 - 1j, -1 + 1j], v=np.ones(5) + 0j)
    assert len(lam) == 3
    for known in (-0.15955391823299256, 0.70313844 + 0.77845926j, 0.70313844 - 0.77845926j):
        assert np.min(np.abs(lam - known)) < 1e-8
    for i in range(3):
        assert np.linalg.norm(o.compute_Mlincomb(dep, lam[i], X[:, i])) < 1e-12


def test_nleigs_scalar():
    """test/nleigs/nleigs_scalar.jl: 0.2 sqrt(l) - 0.6 sin(2l) on [0.01, 4]: one eigenvalue with polynomial interpolation,
    three with the rational one."""
    f = [lambda S: sla.sqrtm(S) if np.ndim(S) == 2 else np.sqrt(S), lambda S: sla.sinm(2 * S) if np.ndim(S) == 2 else np.sin(2 * S)]
    nep = o.SPMF_NEP([np.array([[0.2]]), np.array([[-0.6]])], f)
    Sig = np.array([0.01, 4]) + 0j
    lam, X, _, _ = onl.nleigs(nep, Sig, maxit=100, v=np.ones(1) + 0j, leja=2, isfunm=False)
    _verify(nep, lam, X, 1)
    Xi = -10 ** np.linspace(-6, 5, 10000)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lam, X, _, _ = onl.nleigs(nep, Sig, Xi=Xi, maxit=100, v=np.ones(1) + 0j, leja=2, isfunm=False)
    _verify(nep, lam, X, 3)
    for x in lam:
        assert abs(0.2 * np.sqrt(x) - 0.6 * np.sin(2 * x)) < 1e-9


def test_nleigs_gun_naive_reference_eigenvalue():
    """test/nleigs/nleigs_gun_naive.jl: one eigenvalue in the square 150^2 + 200*[-1-i,..]; it is the gun reference
    eigenvalue of test/gun_native.jl:9.  Also pins the committed golden fixture the GPU tests compare with."""
    import json
    import os
    nep = o.nep_gallery("nlevp_native_gun")
    sq = np.array([-1 - 1j, -1 + 1j, 1 + 1j, 1 - 1j])
    lam, X, res, det = onl.nleigs(nep, 150.0 ** 2 + 200.0 * sq, v=np.ones(nep.n) + 0j)
    assert len(lam) == 1
    assert abs(lam[0] - (22345.116783765 + 0.644998598j)) < 1e-8 * abs(lam[0])
    assert res[0] < 1e-10
    with open(os.path.join(os.path.dirname(__file__), "golden", "nleigs_gun.json")) as f:
        gold = json.load(f)["naive"]
    assert det["kconv"] == gold["kconv"] and det["iterations"] == gold["iterations"]
    assert abs(lam[0] - complex(*gold["lam"][0])) < 1e-10 * abs(lam[0])


def test_lowrank_branch_matches_the_reference_counts_and_the_fullrank_run():
    """The low-rank branches of nleigs (method_nleigs.jl:380-518 with `P.is_low_rank`, rk_helper/rk_nep.jl:43-152) on
    SumNEP(PEP([K, M]), LowRankFactorizedNEP([c1, c2])) as built in test/rk_helper/gun_test_utils.jl:36-43: the golden file holds
    the oracle runs of the variants R1 / R2 / S, for which the reference asserts 21 eigenvalues each
    (test/nleigs/nleigs_gun_variant_{r1,r2,s}.jl); the eigenvalues equal those of the full-rank runs.  Variant R1 is re-run here
    (~20 s), and `backslash_ref` is checked against the full-rank `backslash` on a full-rank operator."""
    import json
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    from make_nleigs_lowrank_golden import gun_lowrank_nep, run
    low = json.load(open(os.path.join(here, "golden", "nleigs_gun_lowrank.json")))
    full = json.load(open(os.path.join(here, "golden", "nleigs_gun.json")))
    for var in ("R1", "R2", "S"):
        assert low[var]["count"] == 21 and max(low[var]["res"]) < 1e-10
    for var in ("R2", "S"):
        a = np.array([complex(*x) for x in low[var]["lam"]])
        b = np.array([complex(*x) for x in full[var]["lam"]])
        assert max(np.min(np.abs(a - x)) / abs(x) for x in b) < 1e-7
    nep, _ = gun_lowrank_nep()
    P = onl.RKNEP(nep)
    assert P.is_low_rank and (P.p, P.q, P.r) == (1, 2, 84)  # ranks 19 + 65 of W1, W2
    for A, L, U in zip(o.get_Av(nep.nep2), nep.nep2.L, nep.nep2.U):
        assert abs(L @ U.T - A).max() < 1e-14
    r1 = run("R1")
    assert r1["count"] == 21
    a = np.array([complex(*x) for x in r1["lam"]])
    b = np.array([complex(*x) for x in low["R1"]["lam"]])
    assert np.max(np.abs(a - b) / np.abs(b)) < 1e-9
