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
