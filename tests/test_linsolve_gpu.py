"""Device LU + triangular solves behind the LinSolver plugin interface, against the CPU oracle (SuperLU).
Mirrors test/linsolver.jl:11-94 and test/rk_helper/cached_lin_solver.jl:8-37 of the reference."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

import nepb200
from nepb200 import B200SPMF, ONE, IDENTITY, PowShift, Monomial, Exp
from oracle import gallery as g
from oracle import nep as o

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def gun_pair():
    K, M, W1, W2 = g.load_gun_matrices()
    onep = o.nep_gallery("nlevp_native_gun")
    dnep = B200SPMF([K, -M, W1, W2], [ONE, IDENTITY, PowShift(0.5, 0.0, 1j), PowShift(0.5, 108.8774 ** 2, 1j)])
    return onep, dnep


def test_gun_factorize_and_backslash_agree_with_oracle():
    # test/linsolver.jl:11-68: lambda = 250^2 + 1im, x = ones(n)
    onep, dnep = gun_pair()
    lam = 250.0 ** 2 + 1j
    Mo = sp.csc_matrix(o.compute_Mder(onep, lam))
    n = dnep.n
    b = np.ones(n, dtype=complex)
    xo = sla.splu(Mo).solve(b)
    norm1 = abs(Mo).sum(axis=0).max()
    fs = nepb200.B200LinSolverCreator().create_linsolver(dnep, lam)
    bs = nepb200.B200BackslashLinSolverCreator().create_linsolver(dnep, lam)
    x1 = fs.lin_solve(b)
    x2 = bs.lin_solve(b)
    for x in (x1, x2):
        assert np.linalg.norm(Mo @ x - b) / norm1 < EPS  # the reference's own criterion
        assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-10
    assert fs.status["flags"] == 0 and fs.status["nperturbed"] == 0
    # matrix right-hand side (method_beyncontour.jl:90-93)
    rng = np.random.default_rng(0)
    B = rng.standard_normal((n, 20)) + 1j * rng.standard_normal((n, 20))
    X = fs.lin_solve(B)
    assert X.shape == (n, 20)
    assert np.linalg.norm(Mo @ X - B) / (norm1 * np.linalg.norm(X)) < 10 * EPS
    assert np.linalg.norm(X - sla.splu(Mo).solve(B)) / np.linalg.norm(X) < 1e-10


def test_refinement_steps():
    # test/linsolver.jl:77-94: refinements 0 vs default
    onep, dnep = gun_pair()
    lam = 150.0 ** 2 + 2j
    Mo = sp.csc_matrix(o.compute_Mder(onep, lam))
    b = np.arange(1.0, dnep.n + 1) + 0j
    s0 = nepb200.B200FactorizeLinSolver(dnep, lam, umfpack_refinements=0)
    s10 = nepb200.B200FactorizeLinSolver(dnep, lam, umfpack_refinements=10)
    x0, x10 = s0.lin_solve(b), s10.lin_solve(b)
    r0 = np.linalg.norm(Mo @ x0 - b) / np.linalg.norm(b)
    r10 = np.linalg.norm(Mo @ x10 - b) / np.linalg.norm(b)
    assert r0 < 1e-10 and r10 <= r0 * 1.5
    assert s10.lu.last_berr < 1e-15


def test_batched_shifts_match_single():
    onep, dnep = gun_pair()
    lams = 150.0 ** 2 + 500.0 * np.exp(2j * np.pi * np.arange(5) / 5)
    lu = nepb200.B200LU(dnep, lams)
    rng = np.random.default_rng(1)
    B = rng.standard_normal((dnep.n, 3)) + 0j
    for s, lam in enumerate(lams):
        Mo = sp.csc_matrix(o.compute_Mder(onep, lam))
        X = lu.solve(B, shift=s)
        assert np.linalg.norm(Mo @ X - B) / (abs(Mo).sum(axis=0).max() * np.linalg.norm(X)) < 10 * EPS
        single = nepb200.B200LU(dnep, [lam])
        assert np.array_equal(single.solve(B), X)  # batch membership must not change a single bit


def test_dense_dep0_and_pivoting():
    # config C1: dense 5x5 (one front, full partial pivoting) and a matrix that needs row exchanges
    A0, A1, tauv = g.dep0_matrices(5)
    dnep = B200SPMF.from_nep(nepb200.DEP([A0, A1], tauv))
    onep = o.nep_gallery("dep0")
    lam = -0.2 + 0.1j
    Mo = o.compute_Mder(onep, lam)
    b = np.ones(5)
    x = nepb200.B200LinSolverCreator().create_linsolver(dnep, lam).lin_solve(b)
    assert np.linalg.norm(Mo @ x - b) < 1e-14
    P = sp.csc_matrix(np.array([[0.0, 2.0, 0.0], [1.0, 0.0, 3.0], [0.0, 4.0, 1e-3]]))
    d = B200SPMF([P], [ONE])
    x = nepb200.B200FactorizeLinSolver(d, 0.0).lin_solve(np.array([1.0, 2.0, 3.0]))
    assert np.linalg.norm(P @ x - [1.0, 2.0, 3.0]) < 1e-15 * np.linalg.norm(P.toarray()) * np.linalg.norm(x)  # |x| ~ 3e3


def test_qdep0_unsymmetric_sparse():
    A0, A1 = g.load_qdep0_matrices()
    n = A0.shape[0]
    onep = o.nep_gallery("qdep0")
    dnep = B200SPMF([-sp.identity(n, format="csc"), A0, A1], [Monomial(2), ONE, Exp(-1.0)])
    lam = -1.0 + 0.2j
    Mo = sp.csc_matrix(o.compute_Mder(onep, lam))
    rng = np.random.default_rng(2)
    B = rng.standard_normal((n, 4)) + 1j * rng.standard_normal((n, 4))
    X = nepb200.B200FactorizeLinSolver(dnep, lam).lin_solve(B)
    assert np.linalg.norm(Mo @ X - B) / (abs(Mo).sum(axis=0).max() * np.linalg.norm(X)) < 10 * EPS
    assert np.linalg.norm(X - sla.splu(Mo).solve(B)) / np.linalg.norm(X) < 1e-9


def test_stencil_pep_solve():
    from nepb200 import synthetic
    mats, _ = synthetic.stencil_pep(48)
    dnep = B200SPMF([m.tocsc() for m in mats], [Monomial(i) for i in range(4)])
    lam = 0.3 + 0.2j
    Mo = sum(m * lam ** i for i, m in enumerate(mats)).tocsc()
    b = np.ones(dnep.n) + 0j
    x = nepb200.B200FactorizeLinSolver(dnep, lam).lin_solve(b)
    assert np.linalg.norm(Mo @ x - b) / (abs(Mo).sum(axis=0).max() * np.linalg.norm(x)) < 10 * EPS


def test_singular_matrix_is_reported():
    """An exactly singular matrix: `factorize` throws SingularException in the reference (UMFPACK, LinSolvers.jl:116); here the zero
    pivot survives the static-pivoting fallback and the solver refuses to be built; the raw handle reports the flags."""
    Z = sp.csc_matrix(np.array([[1.0, 2.0], [2.0, 4.0]]))
    d = B200SPMF([Z], [ONE])
    with pytest.raises(nepb200.SingularException):
        nepb200.B200FactorizeLinSolver(d, 0.0, umfpack_refinements=0)
    lu = nepb200.B200LU(d, [0.0])
    st = lu.status(0)
    assert st["flags"] & 1 and st["flags"] & 8  # zero pivot, row matching was tried
    with pytest.raises(nepb200.SingularException):  # a solve with replaced pivots is verified and rejected, never silent
        lu.solve(np.array([1.0, 0.0]), 0, 2)
    # structurally singular: no row matching exists
    d0 = B200SPMF([sp.csc_matrix(np.array([[1.0, 1.0], [0.0, 0.0]]))], [ONE])
    with pytest.raises(nepb200.SingularException):
        nepb200.B200LU(d0, [0.0])


@pytest.mark.parametrize("transpose", [False, True])
def test_qdep0_sigma0_static_pivoting(transpose):
    """qdep0 at sigma = 0, the configuration of test/infbilanczos.jl:11-15 (operator and its transpose): M(0) = A0 + A1 has 20
    nonzero diagonal entries out of 1000, so pivoting inside the fronts' pivot blocks meets zero pivots; the library falls
    back to the maximum-product row matching + scaling (flag bit 3) and must then agree with SuperLU's partial pivoting."""
    A0, A1 = g.load_qdep0_matrices()
    if transpose:
        A0, A1 = sp.csc_matrix(A0.T), sp.csc_matrix(A1.T)
    n = A0.shape[0]
    dnep = B200SPMF([-sp.identity(n, format="csc"), A0, A1], [Monomial(2), ONE, Exp(-1.0)])
    Mo = sp.csc_matrix(A0 + A1).astype(complex)
    rng = np.random.default_rng(5)
    B = rng.standard_normal((n, 3)) + 1j * rng.standard_normal((n, 3))
    s = nepb200.B200FactorizeLinSolver(dnep, 0.0)
    assert s.status["flags"] == 8 and s.status["nperturbed"] == 0
    X = s.lin_solve(B)
    assert s.lu.last_berr < 1e-15
    assert np.linalg.norm(Mo @ X - B) / (abs(Mo).sum(axis=0).max() * np.linalg.norm(X)) < 10 * EPS
    assert np.linalg.norm(X - sla.splu(Mo).solve(B)) / np.linalg.norm(X) < 1e-9
    # the operator now prefers the matched analysis: other shifts work through it as well, and without refinement
    lam = -1.0 + 0.2j
    M1 = sp.csc_matrix(-lam ** 2 * sp.identity(n) + A0 + np.exp(-lam) * A1)
    s1 = nepb200.B200FactorizeLinSolver(dnep, lam, umfpack_refinements=0)
    assert s1.status["flags"] == 8
    X1 = s1.lin_solve(B)
    assert np.linalg.norm(M1 @ X1 - B) / (abs(M1).sum(axis=0).max() * np.linalg.norm(X1)) < 1e-12
    # device-resident solve with refinement (nepb_lu_solve_block_ex)
    from nepb200 import Block
    from nepb200.dense import solve_block
    Bb, Xb = Block.from_host(B), Block(n, 3)
    solve_block(s, Bb, 0, 3, Xb, 0, alpha=-1.0)
    assert np.linalg.norm(Xb.download() + X) / np.linalg.norm(X) < 1e-12


def test_contour_nodes_with_zero_diagonal():
    """The contour pipeline takes the same fallback: moments of qdep0 on a small circle around 0 (|lambda|^2 ~ 1e-4 on the
    diagonal) against a SuperLU loop."""
    from nepb200.solvers import ContourIntegrator
    A0, A1 = g.load_qdep0_matrices()
    n = A0.shape[0]
    dnep = B200SPMF([-sp.identity(n, format="csc"), A0, A1], [Monomial(2), ONE, Exp(-1.0)])
    N, k = 8, 3
    lams = 0.01 * np.exp(2j * np.pi * (np.arange(N) + 0.5) / N)
    W = np.stack([np.ones(N), lams], axis=1).astype(complex) / N
    Vh = np.random.default_rng(3).standard_normal((n, k))
    integ = ContourIntegrator(dnep, k, 2, N)
    S, flags = integ.integrate(lams, W, Vh, reduce=False)
    integ.close()
    assert np.all(flags & 8) and not np.any(flags & 7)
    ref = np.zeros((n, k, 2), dtype=complex)
    for i, lam in enumerate(lams):
        X = sla.splu(sp.csc_matrix(-lam ** 2 * sp.identity(n) + A0 + np.exp(-lam) * A1)).solve(Vh.astype(complex))
        for j in range(2):
            ref[:, :, j] += W[i, j] * X
    assert np.linalg.norm(S - ref) / np.linalg.norm(ref) < 1e-10


def test_creator_cache_semantics():
    # LinSolverCreators.jl:62-122 and test/rk_helper/cached_lin_solver.jl
    onep, dnep = gun_pair()
    c = nepb200.B200LinSolverCreator(max_factorizations=2)
    a = c.create_linsolver(dnep, 1e4 + 1j)
    assert c.create_linsolver(dnep, 1e4 + 1j) is a
    c.create_linsolver(dnep, 2e4)
    c.create_linsolver(dnep, 3e4)
    assert len(c.recycled_factorizations) == 2
    cache = nepb200.LinSolverCache(dnep)
    y = np.ones(dnep.n, dtype=complex)
    x = cache.solve(2e4 + 5j, y)
    assert len(cache.solvers) == 1
    cache.solve(3e4, y, add_to_cache=False)
    assert len(cache.solvers) == 1
    Mo = sp.csc_matrix(o.compute_Mder(onep, 2e4 + 5j))
    assert np.linalg.norm(Mo @ x - y) / np.linalg.norm(y) < 1e-10
    with pytest.raises(ValueError):
        nepb200.B200LinSolverCreator(precomp_values=[1.0])


def test_symbolic_from_handle_matches_host_analysis():
    onep, dnep = gun_pair()
    perm, parent, sn_ptr, sn_parent = nepb200.symbolic_get(dnep)
    K, M, W1, W2 = g.load_gun_matrices()
    A = (abs(K) + abs(M) + abs(W1) + abs(W2)).tocsc()
    perm2, parent2, _, st = nepb200.analyse_pattern(A)
    assert np.array_equal(perm, perm2) and np.array_equal(parent, parent2)
    assert sn_ptr[0] == 0 and sn_ptr[-1] == dnep.n and np.all(np.diff(sn_ptr) > 0)
    assert np.all((sn_parent == -1) | (sn_parent > np.arange(len(sn_parent))))


def test_gmres_linsolver_on_the_device_operator():
    """GMRESLinSolver (LinSolvers.jl:171-188, creator LinSolverCreators.jl:124-145) on the SPMF of test/newlinsolve.jl:5-12
    (tridiagonal A, B = I, C = diag((1:n)/n), f = 1, s, exp(s); lambda0 = -1.02; preconditioner Pl = Diagonal(M(lambda0)),
    tol = 1e-6): every matrix-vector product is a device SpMM through compute_Mlincomb; the result must agree with the
    device LU to the GMRES tolerance, vector right-hand sides only."""
    n, al, lam0 = 100, 0.01, -1.02
    A = sp.diags([np.ones(n), al * np.ones(n - 1), al * np.ones(n - 1)], [0, 1, -1], format="csc")
    B = sp.identity(n, format="csc")
    Cm = sp.diags(np.arange(1, n + 1) / n, format="csc")
    dnep = B200SPMF([A, B, Cm], [ONE, IDENTITY, Exp(1.0)])
    D0 = sp.diags(dnep.compute_Mder(lam0).diagonal())
    creator = nepb200.GMRESLinSolverCreator(Pl=D0, tol=1e-6, log=True)
    solver = creator.create_linsolver(dnep, lam0)
    assert isinstance(solver, nepb200.GMRESLinSolver)
    b = np.ones(n, dtype=complex)
    x = solver.lin_solve(b)
    xd = nepb200.B200FactorizeLinSolver(dnep, lam0).lin_solve(b)
    Mo = (A + lam0 * B + np.exp(lam0) * Cm).tocsc()
    assert np.linalg.norm(D0.power(-1) @ (Mo @ x - b)) <= 1e-6 * np.linalg.norm(D0.power(-1) @ b) * 1.001
    assert np.linalg.norm(x - xd) / np.linalg.norm(xd) < 1e-4
    tight = nepb200.GMRESLinSolverCreator(Pl=D0).create_linsolver(dnep, lam0).lin_solve(b, tol=1e-12)
    assert np.linalg.norm(tight - xd) / np.linalg.norm(xd) < 1e-9
    with pytest.raises(TypeError):
        solver.lin_solve(np.ones((n, 2)))


def test_deflated_nep_linsolver_on_the_device():
    """DeflatedNEPLinSolver (LinSolvers.jl:209-252) recycling the device factorisation, and the deflated NEP's compute functions
    (nep_deflation.jl:65-197, Generic formulation) over the device operator, on qdep0 (n = 1000): the known eigenvalue
    -1.002466988585764 (errmeasure.jl:55-70) is deflated; compute_Mder / compute_Mlincomb / the Schur-complement solve agree with
    the oracle's MM formulation and a dense solve; resinv on the deflated problem then finds a DIFFERENT eigenpair of qdep0."""
    A0, A1 = g.load_qdep0_matrices()
    n = A0.shape[0]
    onep = o.nep_gallery("qdep0")
    dnep = B200SPMF([-sp.identity(n, format="csc"), A0, A1], [Monomial(2), ONE, Exp(-1.0)])
    lam, v = nepb200.resinv(dnep, lam=-1.0, v=np.ones(n), tol=1e-12, maxit=100, errmeasure=nepb200.ResidualErrmeasure(dnep))
    assert abs(lam - (-1.002466988585764)) < 1e-9
    v = v / np.linalg.norm(v)
    dn = nepb200.deflate_eigpair(dnep, lam, v)
    dn_o = o.deflate_eigpair(onep, lam, v)
    assert dn.n == n + 1
    l2 = -0.9 + 0.3j
    Mo = o.deflated_compute_Mder(dn_o, l2, 0)
    Md = dn.compute_Mder(l2, 0)
    assert np.linalg.norm(np.asarray(Md.todense()) - Mo) <= 1e-10 * np.linalg.norm(Mo)
    rng = np.random.default_rng(4)
    X = rng.standard_normal((n + 1, 2)) + 1j * rng.standard_normal((n + 1, 2))
    zo = o.deflated_compute_Mlincomb(dn_o, l2, X, np.array([1.0, 0.5]))
    assert np.linalg.norm(dn.compute_Mlincomb(l2, X, np.array([1.0, 0.5])) - zo) <= 1e-10 * np.linalg.norm(zo)
    b = rng.standard_normal(n + 1) + 1j * rng.standard_normal(n + 1)
    solver = nepb200.DeflatedNEPLinSolverCreator().create_linsolver(dn, l2)
    x = solver.lin_solve(b)
    xo = np.linalg.solve(Mo, b)
    assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)
    lam2, v2 = nepb200.resinv(dn, lam=-1.0, v=np.ones(n + 1), tol=1e-10, maxit=300,
                              linsolvercreator=nepb200.DeflatedNEPLinSolverCreator(), errmeasure=nepb200.ResidualErrmeasure(dn))
    assert abs(lam2 - lam) > 1e-3  # the deflated pair is not found again
    dn2 = nepb200.deflate_eigpair(dn, lam2, v2 / np.linalg.norm(v2))
    lams, V = nepb200.get_deflated_eigpairs(dn2)
    for l, q in zip(lams, V.T):
        assert np.linalg.norm(o.compute_Mlincomb(onep, l, q / np.linalg.norm(q))) < 1e-6
