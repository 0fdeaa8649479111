"""Proj_SPMF_NEP (src/NEPTypes.jl:652-800), SURVEY.md 8(f) rank 2: the oracle restatement and the product's projection
(all A_i V from one fused GENERAL-mode product with selector blocks) compared on the CPU through the NumPy stand-in operator.
The pep0 docstring numbers (:711-720, :758-770) belong to an older random generator; their structure is what is checked:
W = V = ones gives the sum of all entries, and expanding a projection equals setting the larger one."""
import numpy as np
import scipy.sparse as sp

import nepb200
from host_standin import HostOperator
from oracle import nep as o


def test_oracle_projection_identities():
    nep = o.nep_gallery("pep0")
    n = nep.n
    p = o.create_proj_NEP(nep)
    p.set_projectmatrices(np.ones((n, 1)), np.ones((n, 1)))
    assert abs(o.compute_Mder(p.nep_proj, 0)[0, 0] - np.sum(o.compute_Mder(nep, 0))) < 1e-10
    V = np.eye(n)[:, :2]
    p.set_projectmatrices(V, V)
    assert np.allclose(o.compute_Mder(p.nep_proj, 0), o.compute_Mder(nep, 0)[:2, :2])
    Vn = np.column_stack([V, np.ones(n)])
    p.expand_projectmatrices(Vn, Vn)
    q = o.create_proj_NEP(nep)
    q.set_projectmatrices(Vn, Vn)
    for lam in (0.0, 0.3 - 0.2j):
        assert np.allclose(o.compute_Mder(p.nep_proj, lam), o.compute_Mder(q.nep_proj, lam), rtol=1e-13, atol=1e-12)
        assert np.allclose(o.compute_Mder(q.nep_proj, lam), Vn.conj().T @ o.compute_Mder(nep, lam) @ Vn, rtol=1e-12, atol=1e-10)


def test_product_projection_matches_oracle_with_standin_operator():
    rng = np.random.default_rng(2)
    n, k = 300, 5
    A = [sp.random(n, n, 0.03, random_state=i, format="csc") + (1j * sp.random(n, n, 0.01, random_state=9 + i, format="csc") if i == 2 else 0)
         for i in range(4)]
    A = [sp.csc_matrix(a) for a in A]
    onep = o.SPMF_NEP(A, [o.f_one, o.f_id, o.f_exp(-0.7), o.f_pow(2)])
    fi = [nepb200.ONE, nepb200.IDENTITY, nepb200.Exp(-0.7), nepb200.Monomial(2)]
    op = HostOperator(A, fi)
    W = rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k))
    V = rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k))
    po = o.create_proj_NEP(onep)
    pp = nepb200.create_proj_NEP(op)
    po.set_projectmatrices(W[:, :k - 1], V[:, :k - 1])
    pp.set_projectmatrices(W[:, :k - 1], V[:, :k - 1])
    for Bo, Bp in zip(po.B, pp.B):
        assert np.allclose(Bo, Bp, rtol=1e-13, atol=1e-13)
    po.expand_projectmatrices(W, V)
    pp.expand_projectmatrices(W, V)
    full = nepb200.create_proj_NEP(op).set_projectmatrices(W, V)
    for Bo, Bp, Bf in zip(po.B, pp.B, full.B):
        assert np.allclose(Bo, Bp, rtol=1e-13, atol=1e-13) and np.allclose(Bf, Bp, rtol=1e-13, atol=1e-13)
    lam = 0.4 + 0.1j
    assert np.allclose(pp.compute_Mder(lam), o.compute_Mder(po.nep_proj, lam), rtol=1e-12, atol=1e-12)
    assert np.allclose(pp.compute_Mder(lam, 1), o.compute_Mder(po.nep_proj, lam, 1), rtol=1e-12, atol=1e-12)
    S = rng.standard_normal((3, 3)) * 0.3
    X = rng.standard_normal((k, 3)) + 0j
    assert np.allclose(pp.compute_MM(S, X), o.compute_MM(po.nep_proj, S.astype(complex), X), rtol=1e-11, atol=1e-11)
    a = np.array([1.0, 0.5, 2.0])
    assert np.allclose(pp.compute_Mlincomb(lam, X, a), o.compute_Mlincomb(po.nep_proj, lam, X, a), rtol=1e-11, atol=1e-11)
