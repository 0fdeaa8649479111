"""Deflation (src/nep_deflation.jl) and DeflatedNEPLinSolver (src/LinSolvers.jl:209-252), SURVEY.md 8(f) rank 4: the product's
Generic formulation (binomial expansion, Schur-complement solve) on a NumPy stand-in operator against the oracle's MM formulation
-- the agreement the reference itself asserts in test/deflation.jl:46-92 ("Deflation modes")."""
import numpy as np
import scipy.sparse as sp

import nepb200
from nepb200 import Monomial, ONE, Exp
from oracle import gallery as g
from oracle import nep as o
from oracle import solvers as osol
from host_standin import HostOperator, HostSolverCreator


def _pair():
    A0, A1, tauv = g.dep0_matrices(5)
    onep = o.DEP([A0, A1], [0.0, 0.8])  # test/deflation.jl:9-10
    op = HostOperator([-np.eye(5), A0, A1], [Monomial(1), ONE, Exp(-0.8)])
    return onep, op


def test_deflated_compute_functions_generic_vs_mm():
    onep, op = _pair()
    lam, v = osol.resinv(onep, lam=-0.2, v=np.ones(5), tol=1e-14)
    v = v / np.linalg.norm(v)
    dn_o = o.deflate_eigpair(onep, lam, v)
    dn = nepb200.deflate_eigpair(op, lam, v)
    rng = np.random.default_rng(0)
    for level in range(3):
        # the two formulations differ by M(X, S) (lam I - S)^-1, i.e. by the accuracy of the invariant pair: the first pair is
        # exact to 1e-14, the later ones come from resinv with tol = 1e-10
        rt = 1e-11 if level == 0 else 1e-8
        assert dn.n == dn_o.n == 6 + level
        l2 = 2 + 2j
        for der in range(4):
            Mo = o.deflated_compute_Mder(dn_o, l2, der)
            assert np.linalg.norm(np.asarray(dn.compute_Mder(l2, der)) - Mo) <= rt * max(1.0, np.linalg.norm(Mo))
        X = rng.standard_normal((dn.n, 2)) + 1j * rng.standard_normal((dn.n, 2))
        a = np.array([0.7, -1.3])
        zo = o.deflated_compute_Mlincomb(dn_o, l2, X, a)
        assert np.linalg.norm(dn.compute_Mlincomb(l2, X, a) - zo) <= rt * np.linalg.norm(zo)
        S = np.array([[2, 4], [5, 6.0]])  # test/deflation.jl:68
        MMo = o.deflated_compute_MM(dn_o, S, X)
        assert np.linalg.norm(dn.compute_MM(S, X) - MMo) <= rt * np.linalg.norm(MMo)
        # the Schur-complement solver against a dense solve of the deflated matrix
        b = rng.standard_normal(dn.n) + 1j * rng.standard_normal(dn.n)
        sigma = -0.1 + 0.1j
        x = nepb200.DeflatedNEPLinSolverCreator(HostSolverCreator()).create_linsolver(dn, sigma).lin_solve(b)
        xo = np.linalg.solve(o.deflated_compute_Mder(dn_o, sigma, 0), b)
        assert np.linalg.norm(x - xo) <= 10 * rt * np.linalg.norm(xo)
        # next eigenpair of the deflated problem (resinv on the product's deflated NEP), then deflate again in both
        lam2 = None
        for start in (sigma, 0.5, -0.5 + 1j, 1.0 + 1j, -1.0):  # resinv converges linearly from a fixed shift: try a few
            try:
                lam2, v2 = nepb200.resinv(dn, lam=start, v=np.ones(dn.n), tol=1e-10, maxit=300,
                                          linsolvercreator=nepb200.DeflatedNEPLinSolverCreator(HostSolverCreator()),
                                          errmeasure=nepb200.ResidualErrmeasure(dn))
                break
            except nepb200.NoConvergenceException:
                continue
        assert lam2 is not None
        v2 = v2 / np.linalg.norm(v2)
        dn = nepb200.deflate_eigpair(dn, lam2, v2)
        dn_o = o.deflate_eigpair(dn_o, lam2, v2)
    lams, V = nepb200.get_deflated_eigpairs(dn)
    assert len(lams) == 4 and len(np.unique(np.round(lams, 6))) == 4  # four different eigenvalues were found
    for l, w in zip(lams, V.T):  # test/deflation.jl:34-40: eigenpairs of the ORIGINAL problem
        assert np.linalg.norm(o.compute_Mlincomb(onep, l, w / np.linalg.norm(w))) < np.sqrt(np.finfo(float).eps)
