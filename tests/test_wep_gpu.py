"""WEP-native device path (csrc/wep.cu, nepb200/wep.py) against the oracle (oracle/wep.py) and the reference's own checks
(test/wep_small.jl): format equivalence, compute_Mlincomb with derivatives, the boundary operator, the Schur-complement product
and solver, resinv / quasinewton-style convergence to the reference eigenvalue, iar."""
import numpy as np
import pytest

import nepb200
from nepb200 import B200SPMF, Block
from oracle import nep as o
from oracle import wep as ow

pytestmark = pytest.mark.gpu

LAMREF = -2.743228671961724 - 3.1439375599649972j  # test/wep_small.jl:35


def _rand(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def test_formats_agree_on_the_device():
    """test/wep_small.jl:17-26: the SPMF format (through the fused SPMF kernels) and the native format (stencil + direct
    transforms) give the same M(lambda) * ones to 1e-14; both equal the oracle.  nz = 7 as in the reference (p = 17 terms) and
    nz = 5 (p = 13)."""
    for nx, nz in ((11, 7), (11, 5)):
        _formats(nx, nz)


def _formats(nx, nz):
    spmf = B200SPMF.from_nep(nepb200.nep_gallery_WEP(nx=nx, nz=nz, benchmark_problem="TAUSCH", neptype="SPMF"))
    nep = nepb200.nep_gallery_WEP(nx=nx, nz=nz, benchmark_problem="TAUSCH", neptype="WEP")
    onep = ow.nep_gallery_wep(nx=nx, nz=nz, benchmark_problem="TAUSCH", neptype="WEP")
    lam = -1.3 - 0.31j
    v1 = spmf.compute_Mlincomb(lam, np.ones(nep.n))
    v2 = nep.compute_Mlincomb(lam, np.ones(nep.n))
    vo = o.compute_Mlincomb(onep, lam, np.ones(nep.n))
    assert nep.size() == (nx * nz + 2 * nz,) * 2 and nep.size(1) == nep.n
    assert np.linalg.norm(v1 - v2) / np.linalg.norm(v1) < 1e-14
    assert np.linalg.norm(v2 - vo) / np.linalg.norm(vo) < 1e-14
    with pytest.raises(NotImplementedError):
        nep.compute_Mder(lam)
    with pytest.raises(ValueError):
        nep.compute_Mlincomb(lam, np.ones(nep.n - 1))
    with pytest.raises(ValueError):
        nep.compute_Mlincomb(lam, np.ones((nep.n, 2)), np.ones(3))


@pytest.mark.parametrize("wg,nx,nz,na", [("TAUSCH", 11, 7, 1), ("JARLEBRING", 13, 9, 2), ("JARLEBRING", 109, 105, 3),
                                         ("JARLEBRING", 109, 105, 6), ("TAUSCH", 40, 33, 12), ("JARLEBRING", 64, 128, 4)])
def test_mlincomb_matches_oracle(wg, nx, nz, na):
    """sum_j a_j M^{(j)}(lambda) v_j for 1..12 columns (derivative columns of the interior up to the second, of the boundary
    operator all of them), odd and even nz; 1e-13 relative."""
    nep = nepb200.nep_gallery_WEP(nx=nx, nz=nz, benchmark_problem=wg, neptype="WEP")
    onep = ow.nep_gallery_wep(nx=nx, nz=nz, benchmark_problem=wg, neptype="WEP")
    rng = np.random.default_rng(nx * nz + na)
    V, a = _rand(rng, nep.n, na), _rand(rng, na)
    lam = -2.7 - 3.1j
    z = nep.compute_Mlincomb(lam, V, a)
    zo = ow.wep_compute_Mlincomb(onep, lam, V, a)
    assert np.linalg.norm(z - zo) <= 1e-13 * np.linalg.norm(zo)
    if na >= 2:  # startder (NEPCore.jl:156-160) and a zero coefficient
        z = nep.compute_Mlincomb(lam, V[:, :1], np.array([1.0]), 1)
        zo = ow.wep_compute_Mlincomb(onep, lam, np.column_stack([V[:, 0], V[:, 0]]), np.array([0, 1.0]))
        assert np.linalg.norm(z - zo) <= 1e-13 * np.linalg.norm(zo)
    # bitwise reproducible
    assert np.array_equal(nep.compute_Mlincomb(lam, V, a), nep.compute_Mlincomb(lam, V, a))


def test_boundary_operator_schur_product_and_solver():
    nx, nz = 109, 105
    nep = nepb200.nep_gallery_WEP(nx=nx, nz=nz, benchmark_problem="JARLEBRING", neptype="WEP")
    onep = ow.nep_gallery_wep(nx=nx, nz=nz, benchmark_problem="JARLEBRING", neptype="WEP")
    rng = np.random.default_rng(5)
    lam = -3 - 3.5j
    x = _rand(rng, 2 * nz)
    assert np.linalg.norm(nep.Pinv(lam, x) - onep.Pinv(lam, x)) <= 1e-13 * np.linalg.norm(x)
    q = _rand(rng, nx * nz)
    mv = nepb200.SchurMatVec(nep, lam)
    yo = ow.schur_matvec(onep, lam, q)
    assert np.linalg.norm(mv(q) - yo) <= 1e-13 * np.linalg.norm(yo)
    S = nepb200.construct_WEP_schur_complement(nep, lam)
    So = ow.construct_WEP_schur_complement(onep, lam)
    assert abs(S - So).max() <= 1e-12 * abs(So).max()
    # lin_solve (Ringh, Prop. 2.1) with the device LU of the Schur complement: M(lambda) x = b
    b = _rand(rng, nep.n)
    for kind in ("factorized", "backslash"):
        solver = nepb200.WEPLinSolverCreator(solver_type=kind).create_linsolver(nep, lam)
        xs = solver.lin_solve(b)
        assert np.linalg.norm(nep.compute_Mlincomb(lam, xs) - b) <= 1e-10 * np.linalg.norm(b)
        xo = ow.WEPFactorizedLinSolver(onep, lam).lin_solve(b)
        assert np.linalg.norm(xs - xo) <= 1e-9 * np.linalg.norm(xo)
    with pytest.raises(TypeError):
        nepb200.WEPLinSolverCreator().create_linsolver(object(), lam)
    with pytest.raises(ValueError):
        nepb200.WEPLinSolverCreator(solver_type="nope").create_linsolver(nep, lam)


def test_gmres_on_the_matrix_free_schur_complement_small():
    """WEPGMRESLinSolver (Waveguide.jl:424-456) without a preconditioner on a small grid (full GMRES)."""
    nep = nepb200.nep_gallery_WEP(nx=11, nz=7, benchmark_problem="TAUSCH", neptype="WEP")
    rng = np.random.default_rng(6)
    lam = -1.3 - 0.31j
    b = _rand(rng, nep.n)
    solver = nepb200.WEPLinSolverCreator(solver_type="gmres", kwargs=(("reltol", 1e-12), ("restart", 77), ("maxiter", 77))).create_linsolver(nep, lam)
    x = solver.lin_solve(b)
    assert np.linalg.norm(nep.compute_Mlincomb(lam, x) - b) <= 1e-9 * np.linalg.norm(b)


def test_resinv_and_iar_reach_the_reference_eigenvalue():
    """test/wep_small.jl:28-47 and :62-72: JARLEBRING, nx = 109, nz = 105, lambda0 = -3 - 3.5im, v0 = ones / norm: resinv with
    the (device) Schur-complement solver converges to the reference eigenvalue with residual < 1e-10; iar with sigma = lambda0
    finds it among 3 eigenvalues (min |lambda_ref - lambda| < 1e-10)."""
    nep = nepb200.nep_gallery_WEP(nx=3 * 5 * 7 + 4, nz=3 * 5 * 7, benchmark_problem="JARLEBRING", neptype="WEP")
    n = nep.n
    v0 = np.ones(n) / np.sqrt(n)

    class RefErr:  # EigvalReferenceErrmeasure (errmeasure.jl:150-160)
        def estimate_error(self, lam, v):
            return abs(lam - LAMREF) / abs(lam)
    creator = nepb200.WEPLinSolverCreator(solver_type="factorized")
    lam, v = nepb200.resinv(nep, lam=-3 - 3.5j, v=v0, tol=1e-12, errmeasure=RefErr(), linsolvercreator=creator)
    assert abs(lam - LAMREF) < 1e-11 * abs(lam)
    assert np.linalg.norm(nep.compute_Mlincomb(lam, v)) / np.linalg.norm(v) < 1e-10
    lams, V = nepb200.iar(nep, sigma=-3 - 3.5j, neigs=3, maxit=100, v=v0, tol=1e-8, linsolvercreator=creator)[:2]
    assert len(lams) == 3 and np.min(np.abs(LAMREF - lams)) < 1e-10
    # tiar with the basis in HBM (config C5 of BASELINE.json in small): the same three eigenvalues (iar == tiar, test/tiar.jl)
    lt, Qt = nepb200.tiar_device(nep, sigma=-3 - 3.5j, neigs=3, maxit=100, v=v0, tol=1e-8, linsolvercreator=creator)[:2]
    assert len(lt) == 3 and np.min(np.abs(LAMREF - lt)) < 1e-10
    assert max(np.min(np.abs(lams - x)) for x in lt) < 1e-8
    for i in range(3):
        assert np.linalg.norm(nep.compute_Mlincomb(lt[i], Qt[:, i])) / np.linalg.norm(Qt[:, i]) < 1e-8


def test_resinv_with_gmres_and_the_sylvester_preconditioner():
    """test/wep_small.jl:55-61: resinv with WEPLinSolverCreator(solver_type = :gmres, kwargs = ((:Pl, precond), (:reltol, 1e-7)))
    and precond = wep_generate_preconditioner(nep, 3*7, lambda0).  The Schur-complement products of GMRES run on the device; the
    preconditioner is the oracle's restatement of waveguide_preconditioner.jl, handed over as the `Pl` callable (the product
    ships the solver, not this preconditioner)."""
    nep = nepb200.nep_gallery_WEP(nx=3 * 5 * 7 + 4, nz=3 * 5 * 7, benchmark_problem="JARLEBRING", neptype="WEP")
    onep = ow.nep_gallery_wep(nx=3 * 5 * 7 + 4, nz=3 * 5 * 7, benchmark_problem="JARLEBRING", neptype="WEP")
    lam0 = -3 - 3.5j
    precond = ow.wep_generate_preconditioner(onep, 3 * 7, lam0)
    n = nep.n
    v0 = np.ones(n) / np.sqrt(n)

    class RefErr:
        def estimate_error(self, lam, v):
            return abs(lam - LAMREF) / abs(lam)
    creator = nepb200.WEPLinSolverCreator(solver_type="gmres", kwargs=(("Pl", precond), ("reltol", 1e-7)))
    lam, v = nepb200.resinv(nep, lam=lam0, v=v0, tol=1e-12, errmeasure=RefErr(), linsolvercreator=creator)
    assert abs(lam - LAMREF) < 1e-11 * abs(lam)
    assert np.linalg.norm(nep.compute_Mlincomb(lam, v)) / np.linalg.norm(v) < 1e-10
