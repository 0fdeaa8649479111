"""infbilanczos (src/method_infbilanczos.jl), a "next" row of SURVEY.md 8(f): the oracle restatement pinned to the literal
tridiagonal matrix of test/infbilanczos.jl:19-22, and the product's host recurrences (which turn the reference's O(m^3 n)
double loop of compute_Mlincomb calls into one fused multi-term product per bilinear form) checked against the same literal
with a NumPy stand-in for the device operator -- the ABI calls themselves are covered by the GPU tests."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

import nepb200
from nepb200 import _lib
from oracle import nep as o
from oracle import solvers as s

TSTAR = np.array([[-1.665117675679600, 5.780562035399026, 0, 0],
                  [5.780562035399026, 11.562308485001218, -18.839546184493731, 0],
                  [0, 18.839546184493734, -15.213756300995186, 9.788512505128466],
                  [0, 0, 9.788512505128464, -0.120825360586847]])


def _qdep0_pair():
    nep = o.nep_gallery("qdep0")
    nept = o.SPMF_NEP([sp.csc_matrix(A.T) for A in nep.A], nep.fi)
    return nep, nept


def test_oracle_infbilanczos_reference_literals():
    nep, nept = _qdep0_pair()
    n = nep.n
    kw = dict(sigma=0, v=np.ones(n), u=np.ones(n), check_error_every=3, tol=1e-7, errmeasure=o.residual_errmeasure(nep))
    lam, V, T = s.infbilanczos(nep, nept, maxit=40, neigs=3, **kw)
    n0 = min(4, len(lam))
    assert np.linalg.norm(TSTAR[:n0, :n0] - T[:n0, :n0], 2) < 1e-10  # test/infbilanczos.jl:23-24
    assert sum(np.linalg.norm(o.compute_Mlincomb(nep, lam[i], V[:, i])) < 1e-7 for i in range(len(lam))) == 3
    lam, V, T = s.infbilanczos(nep, nept, maxit=30, neigs=np.inf, **kw)  # :33-39
    assert len(lam) == 3
    for i in range(3):
        assert np.linalg.norm(o.compute_Mlincomb(nep, lam[i], V[:, i])) / np.linalg.norm(V[:, i]) < 1e-6
    with pytest.raises(s.NoConvergenceException):  # :42-46
        s.infbilanczos(nep, nept, maxit=9, neigs=8, **kw)
    # docstring (:17-26): dep0, neigs = 3, residual ~1e-14
    dep = o.nep_gallery("dep0")
    A = o.get_Av(dep)
    dept = o.SPMF_NEP([np.array(a.T) if not sp.issparse(a) else a.T for a in A], o.get_fv(dep))
    lam, V, _ = s.infbilanczos(dep, dept, neigs=3, v=np.ones(5), u=np.ones(5))
    assert len(lam) == 3 and np.linalg.norm(o.compute_Mlincomb(dep, lam[0], V[:, 0])) < 1e-12


from host_standin import HostOperator as _HostOperator, HostSolverCreator as _HostSolverCreator, HostResidual as _HostResidual  # noqa: E402


def test_product_infbilanczos_host_recurrences_match_reference_literal():
    onep, _ = _qdep0_pair()
    fi = [nepb200.Monomial(2), nepb200.ONE, nepb200.Exp(-1.0)]
    op = _HostOperator([sp.csc_matrix(A) for A in onep.A], fi)
    opt = _HostOperator([sp.csc_matrix(A.T) for A in onep.A], fi)
    n = op.n
    kw = dict(sigma=0, v=np.ones(n), u=np.ones(n), check_error_every=3, tol=1e-7, errmeasure=_HostResidual(op),
              linsolvercreator=_HostSolverCreator(), linsolvertcreator=_HostSolverCreator())
    lam, V, T = nepb200.infbilanczos(op, opt, maxit=40, neigs=3, **kw)
    n0 = min(4, len(lam))
    assert np.linalg.norm(TSTAR[:n0, :n0] - T[:n0, :n0], 2) < 1e-10
    lo, Vo, To = s.infbilanczos(*_qdep0_pair(), maxit=40, neigs=3, sigma=0, v=np.ones(n), u=np.ones(n), check_error_every=3, tol=1e-7,
                                errmeasure=o.residual_errmeasure(onep))
    # a Lanczos recurrence: rounding differences grow along the tridiagonal (1e-5 at entry 30), the leading block is stable
    assert T.shape == To.shape and np.abs(T[:10, :10] - To[:10, :10]).max() < 1e-9 * np.abs(To).max()
    assert np.abs(T - To).max() < 1e-3 * np.abs(To).max()
    assert len(lam) == len(lo) == 3
    for x in lam:
        assert np.min(np.abs(lo - x)) < 1e-9
    with pytest.raises(nepb200.NoConvergenceException):
        nepb200.infbilanczos(op, opt, maxit=9, neigs=8, **kw)
