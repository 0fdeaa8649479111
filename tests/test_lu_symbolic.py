"""Integer work of the LU analysis (ordering, elimination tree, column counts) -- host only, bit-exact against
an independent restatement (dense boolean elimination / Liu's algorithm in Python)."""
import numpy as np
import scipy.sparse as sp

import nepb200
from oracle import gallery as g


def _etree_counts_dense(P):
    """P: boolean symmetric pattern (permuted).  Returns etree parent and column counts by explicit elimination."""
    n = P.shape[0]
    L = np.tril(P | np.eye(n, dtype=bool))
    for j in range(n):
        r = np.nonzero(L[j + 1:, j])[0] + j + 1
        if len(r):
            L[np.ix_(r, r)] |= np.tril(np.ones((len(r), len(r)), dtype=bool))
    parent = np.full(n, -1, dtype=np.int32)
    for j in range(n):
        r = np.nonzero(L[j + 1:, j])[0]
        if len(r):
            parent[j] = r[0] + j + 1
    return parent, L.sum(axis=0).astype(np.int32)


def test_small_patterns_against_dense_elimination():
    rng = np.random.default_rng(0)
    for n, dens in ((1, 1.0), (7, 0.3), (40, 0.08), (120, 0.03), (200, 0.02)):
        A = sp.random(n, n, dens, random_state=int(rng.integers(1 << 30)), format="csc") + sp.identity(n, format="csc")
        for ordering in (0, 1):
            perm, parent, cc, st = nepb200.analyse_pattern(A, ordering=ordering)
            assert sorted(perm.tolist()) == list(range(n))
            S = ((abs(A) + abs(A.T)) != 0).toarray()
            P = S[np.ix_(perm, perm)]
            pref, ccref = _etree_counts_dense(P)
            assert np.array_equal(parent, pref)
            assert np.array_equal(cc, ccref)
            assert all(parent[j] == -1 or parent[j] > j for j in range(n))  # postordered
            assert st["nnz_factor"] >= 2 * cc.sum() - n  # supernodes only add explicit zeros


def test_ordering_reduces_fill_on_gun():
    K, M, W1, W2 = g.load_gun_matrices()
    A = (abs(K) + abs(M) + abs(W1) + abs(W2)).tocsc()
    perm, parent, cc, st = nepb200.analyse_pattern(A)
    _, _, ccn, stn = nepb200.analyse_pattern(A, ordering=1)
    assert sorted(perm.tolist()) == list(range(A.shape[0]))
    assert cc.sum() < 1.6e6 < 3.9e6 < ccn.sum()  # AMD: 1.48 M entries in L; natural order: 4.0 M
    # determinism: the analysis is pure integer work
    perm2, parent2, cc2, _ = nepb200.analyse_pattern(A)
    assert np.array_equal(perm, perm2) and np.array_equal(parent, parent2) and np.array_equal(cc, cc2)


def test_grid_stencil_pattern():
    from nepb200 import synthetic
    mats, _ = synthetic.stencil_pep(40)
    perm, parent, cc, st = nepb200.analyse_pattern(mats[0])
    n = 1600
    assert sorted(perm.tolist()) == list(range(n))
    S = (mats[0] != 0).toarray()
    P = S[np.ix_(perm, perm)]
    pref, ccref = _etree_counts_dense(P)
    assert np.array_equal(parent, pref) and np.array_equal(cc, ccref)


def test_fast_and_legacy_symbolic_algorithms_agree(monkeypatch):
    """Column counts (Gilbert-Ng-Peyton skeleton counting) and front row structures (supernodal merge of the children's update
    rows) replace the O(nnz(L)) row-subtree traversals; NEPB_LU_CHECK=1 makes the library compute both variants and fail unless
    the column counts and every front's row list are identical.  Also the legacy variant alone (NEPB_LU_LEGACY=1) must give
    the same public results."""
    from nepb200 import synthetic
    rng = np.random.default_rng(1)
    pats = [sp.random(n, n, d, random_state=int(rng.integers(1 << 30)), format="csc") + sp.identity(n, format="csc")
            for n, d in ((1, 1.0), (2, 1.0), (9, 0.3), (60, 0.06), (300, 0.01), (2500, 0.0015), (3000, 0.0002))]
    K, M, W1, W2 = g.load_gun_matrices()
    pats.append((abs(K) + abs(M) + abs(W1) + abs(W2)).tocsc())
    indptr, indices, _ = synthetic.stencil_pattern(120)
    pats.append(sp.csr_matrix((np.ones(len(indices)), indices, indptr), shape=(120 * 120, 120 * 120)).tocsc())
    for A in pats:
        for ordering in (0, 1):
            for relax, mx in ((0, 0), (4, 8), (1, 1)):
                monkeypatch.delenv("NEPB_LU_LEGACY", raising=False)
                monkeypatch.setenv("NEPB_LU_CHECK", "1")
                perm, parent, cc, st = nepb200.analyse_pattern(A, ordering=ordering, relax_leaf=relax, max_np=mx, fronts=True)
                monkeypatch.delenv("NEPB_LU_CHECK")
                monkeypatch.setenv("NEPB_LU_LEGACY", "1")
                perm2, parent2, cc2, st2 = nepb200.analyse_pattern(A, ordering=ordering, relax_leaf=relax, max_np=mx, fronts=True)
                assert np.array_equal(perm, perm2) and np.array_equal(parent, parent2) and np.array_equal(cc, cc2)
                for key in ("nnz_factor", "front_entries", "nfronts", "nlevels", "max_front", "flops", "solve_rows"):
                    assert st[key] == st2[key]
                assert np.array_equal(st["np"], st2["np"]) and np.array_equal(st["nf"], st2["nf"]) and np.array_equal(st["level"], st2["level"])


def test_max_product_matching_is_optimal_and_scaled():
    """nepb_lu_matching (lu_matching.cpp): the static-pivoting preprocessing that stands in for UMFPACK's numerical pivoting.
    Integer result (a permutation) checked for optimality against SciPy's sparse assignment solver; the scalings must turn
    the matrix into an I-matrix (matched entries 1, all others <= 1)."""
    from scipy.sparse.csgraph import min_weight_full_bipartite_matching
    rng = np.random.default_rng(11)
    A0, A1 = g.load_qdep0_matrices()
    cases = [sp.csc_matrix(A0 + A1), sp.csc_matrix((A0 + A1).T)]  # test/infbilanczos.jl at sigma = 0: 980 zero diagonal entries
    for n, dens in ((1, 1.0), (6, 0.5), (60, 0.1), (300, 0.02)):
        # a hidden permutation guarantees structural non-singularity; values over ten orders of magnitude
        R = sp.random(n, n, dens, random_state=int(rng.integers(1 << 30)), format="csr")
        Pm = sp.csr_matrix((np.ones(n), (np.arange(n), rng.permutation(n))), shape=(n, n))
        M = (R + Pm).tocsc()
        M.data = np.exp(rng.uniform(-12, 12, M.nnz)) * rng.choice([-1.0, 1.0], M.nnz)
        cases.append(M)
    for M in cases:
        n = M.shape[0]
        roc, dr, dc = nepb200.matching(M)
        assert sorted(roc.tolist()) == list(range(n))
        Aabs = abs(M).tocsr()
        matched = np.asarray(Aabs[roc, np.arange(n)]).ravel()
        assert np.all(matched > 0)
        W = Aabs.copy()
        W.data = 40.0 - np.log(W.data)  # positive weights: SciPy's solver drops explicit zeros
        r, c = min_weight_full_bipartite_matching(W)
        best = np.log(np.asarray(Aabs[r, c]).ravel()).sum()
        assert abs(np.log(matched).sum() - best) <= 1e-9 * max(1.0, abs(best))
        S = sp.diags(dr) @ Aabs @ sp.diags(dc)
        assert S.max() <= 1 + 1e-12
        assert np.allclose(np.asarray(S.tocsr()[roc, np.arange(n)]).ravel(), 1.0, rtol=1e-12)
    # structurally singular patterns are reported
    import pytest
    with pytest.raises(nepb200.SingularException):
        nepb200.matching(sp.csc_matrix(np.array([[1.0, 1.0, 0.0], [1.0, 1.0, 0.0], [1.0, 1.0, 0.0]])))
    # static pivoting on the matched, scaled qdep0 matrix is stable without any further row exchange (SuperLU told not to pivot)
    import scipy.sparse.linalg as sla
    M = cases[0]
    roc, dr, dc = nepb200.matching(M)
    B = (sp.diags(dr) @ M @ sp.diags(dc)).tocsr()[roc, :].tocsc()
    lu = sla.splu(B, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    assert np.array_equal(lu.perm_r, lu.perm_c)
    b = np.ones(M.shape[0])
    assert np.linalg.norm(B @ lu.solve(b) - b) / np.linalg.norm(b) < 1e-8
