"""The WEP oracle (oracle/wep.py) against the reference's own tests (test/wep_small.jl)."""
import numpy as np

from oracle import nep as o
from oracle import solvers as osol
from oracle import wep as ow

LAMREF = -2.743228671961724 - 3.1439375599649972j  # test/wep_small.jl:35


def test_spmf_and_native_format_agree():
    """test/wep_small.jl:17-26: nx = 11, nz = 7, TAUSCH, lambda = -1.3 - 0.31im, v = ones: 1e-14."""
    nx, nz = 11, 7
    A, f = ow.nep_gallery_wep(nx=nx, nz=nz, benchmark_problem="TAUSCH", neptype="SPMF")
    nep = ow.nep_gallery_wep(nx=nx, nz=nz, benchmark_problem="TAUSCH", neptype="WEP")
    assert len(A) == 3 + 2 * nz and nep.n == nx * nz + 2 * nz
    lam = -1.3 - 0.31j
    v1 = ow.spmf_compute_Mlincomb(A, f, lam, np.ones(nep.n))
    v2 = o.compute_Mlincomb(nep, lam, np.ones(nep.n))
    assert np.linalg.norm(v1 - v2) / np.linalg.norm(v1) < 1e-14
    # JARLEBRING as well, random vector
    A, f = ow.nep_gallery_wep(nx=13, nz=9, benchmark_problem="JARLEBRING", neptype="SPMF")
    nep = ow.nep_gallery_wep(nx=13, nz=9, benchmark_problem="JARLEBRING", neptype="WEP")
    rng = np.random.default_rng(0)
    v = rng.standard_normal(nep.n) + 1j * rng.standard_normal(nep.n)
    v1 = ow.spmf_compute_Mlincomb(A, f, -3 - 3.5j, v)
    v2 = o.compute_Mlincomb(nep, -3 - 3.5j, v)
    assert np.linalg.norm(v1 - v2) / np.linalg.norm(v1) < 1e-14


def test_derivatives_by_finite_differences_and_schur_solver():
    """compute_Mlincomb with several columns = sum_i a_i M^{(i)} v_i (checked against central differences of M(lambda) v), and
    the Schur-complement solver (Ringh, Prop. 2.1; Waveguide.jl:523-567) inverts M(lambda); SchurMatVec equals the assembled
    Schur complement."""
    nep = ow.nep_gallery_wep(nx=15, nz=9, benchmark_problem="JARLEBRING", neptype="WEP")
    rng = np.random.default_rng(1)
    lam = -2.7 - 3.1j
    v = rng.standard_normal(nep.n) + 1j * rng.standard_normal(nep.n)
    h = 1e-4
    M = lambda l: ow.wep_compute_Mlincomb(nep, l, v)  # noqa: E731
    d1 = (M(lam + h) - M(lam - h)) / (2 * h)
    d2 = (M(lam + h) - 2 * M(lam) + M(lam - h)) / h ** 2
    z1 = ow.wep_compute_Mlincomb(nep, lam, np.column_stack([v, v]), np.array([0, 1.0]))
    z2 = ow.wep_compute_Mlincomb(nep, lam, np.column_stack([v, v, v]), np.array([0, 0, 1.0]))
    assert np.linalg.norm(z1 - d1) < 1e-6 * np.linalg.norm(d1)
    assert np.linalg.norm(z2 - d2) < 1e-4 * np.linalg.norm(d2)
    V = rng.standard_normal((nep.n, 5)) + 1j * rng.standard_normal((nep.n, 5))
    a = rng.standard_normal(5) + 1j * rng.standard_normal(5)
    z = ow.wep_compute_Mlincomb(nep, lam, V, a)
    zs = sum(ow.wep_compute_Mlincomb(nep, lam, np.column_stack([V[:, j]] * (j + 1)), np.eye(j + 1)[j] * a[j]) for j in range(5))
    assert np.linalg.norm(z - zs) < 1e-13 * np.linalg.norm(z)
    solver = ow.WEPFactorizedLinSolver(nep, lam)
    x = solver.lin_solve(v)
    assert np.linalg.norm(ow.wep_compute_Mlincomb(nep, lam, x) - v) < 1e-11 * np.linalg.norm(v)
    q = rng.standard_normal(nep.nx * nep.nz) + 0j
    S = ow.construct_WEP_schur_complement(nep, lam)
    assert np.linalg.norm(S @ q - ow.schur_matvec(nep, lam, q)) < 1e-12 * np.linalg.norm(S @ q)


def test_resinv_reaches_the_reference_eigenvalue():
    """test/wep_small.jl:28-47: JARLEBRING, nx = 109, nz = 105, lambda0 = -3 - 3.5im, v0 = ones: resinv with the Schur-complement
    solver and EigvalReferenceErrmeasure(lambda_ref) at tol 1e-12 converges, residual < 1e-10."""
    nep = ow.nep_gallery_wep(nx=3 * 5 * 7 + 4, nz=3 * 5 * 7, benchmark_problem="JARLEBRING", neptype="WEP")
    n = nep.n
    v0 = np.ones(n) / np.sqrt(n)
    err = lambda lam, v: abs(lam - LAMREF) / abs(lam)  # noqa: E731  (errmeasure.jl:150-160)
    lam, v = osol.resinv(nep, lam=-3 - 3.5j, v=v0, tol=1e-12, errmeasure=err, linsolvercreator=ow.WEPLinSolverCreator())
    assert abs(lam - LAMREF) < 1e-11 * abs(lam)
    assert np.linalg.norm(o.compute_Mlincomb(nep, lam, v)) / np.linalg.norm(v) < 1e-10


def test_sylvester_smw_preconditioner_literal():
    """test/wep_small.jl:28-31: with N = nz domains the Sylvester-SMW preconditioner (waveguide_preconditioner.jl) inverts the
    Schur complement: ldiv!(precond, SchurMatVec * b1) == b1 to 1e-14; and its Sylvester solver solves A X + X B = C."""
    nx, nz = 11, 7
    nep = ow.nep_gallery_wep(nx=nx, nz=nz, benchmark_problem="TAUSCH", neptype="WEP")
    lam = -1.3 - 0.31j
    rng = np.random.default_rng(0)
    C = rng.standard_normal((nz, nx)) + 1j * rng.standard_normal((nz, nx))
    X = ow.solve_wg_sylvester_fft(C, lam, nep.k_bar, nep.hx, nep.hz)
    assert np.linalg.norm(nep.A(lam) @ X + X @ nep.Dxx.toarray() - C) < 1e-13 * np.linalg.norm(C)
    precond = ow.wep_generate_preconditioner(nep, nz, lam)
    b1 = rng.random(nx * nz) + 1j * rng.random(nx * nz)
    b2 = precond.ldiv(ow.schur_matvec(nep, lam, b1))
    assert np.linalg.norm(b1 - b2) / np.linalg.norm(b1) < 1e-14
    import pytest
    with pytest.raises(ValueError):
        ow.wep_generate_preconditioner(ow.nep_gallery_wep(nx=12, nz=7, neptype="WEP"), 7, lam)   # nx != nz + 4
    with pytest.raises(ValueError):
        ow.wep_generate_preconditioner(nep, 3, lam)   # nz / N not an integer
