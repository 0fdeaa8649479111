"""The oracle pinned against the reference's own literal known answers.  CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import gallery as g
from oracle import nep as o


def test_msws_dep0_mder_literal():
    # src/NEPTypes.jl:74-75: compute_Mder(nep_gallery("dep0"), 3.0)[1,1] == -2.942777908030041
    nep = o.nep_gallery("dep0")
    assert o.compute_Mder(nep, 3.0)[0, 0].real == -2.942777908030041


def test_dep0_100_mlincomb_norm_literal():
    # src/Gallery.jl:174-176
    nep = o.nep_gallery("dep0", 100)
    z = o.compute_Mlincomb(nep, 1.0 + 1.0j, np.ones(100))
    assert abs(np.linalg.norm(z) - 57.498446538064954) < 1e-13


def test_dep0_eigenvalue_literal():
    # docs/src/methods.md:18-19
    nep = o.nep_gallery("dep0")
    M = o.compute_Mder(nep, -0.15955391823299256)
    assert np.linalg.svd(M, compute_uv=False)[-1] < 1e-14


def test_gun_one_norms_literal():
    # test/rk_helper/gun_test_utils.jl:50-53
    K, M, W1, W2 = g.load_gun_matrices()
    lit = [1.474544889815002e+05, 2.726114618171165e-02, 2.328612251920476e+00, 3.793375498194695e+00]
    for A, ref in zip((K, M, W1, W2), lit):
        assert abs(abs(A).sum(axis=0).max() - ref) <= 1e-15 * ref * 4
    assert K.shape == (9956, 9956)
    assert (K.nnz, M.nnz, W1.nnz, W2.nnz) == (148308, 148318, 57, 293)


def test_gun_reference_eigenvalue_literal():
    # test/gun_native.jl:9 -- M(lambda_ref) is numerically singular
    import scipy.sparse.linalg as sla
    nep = o.nep_gallery("nlevp_native_gun")
    lam = 22345.116783765 + 0.644998598j
    M = sp.csc_matrix(o.compute_Mder(nep, lam))
    lu = sla.splu(M)
    # inverse iteration: ||M x|| / ||x|| after two steps ~ smallest singular value
    x = np.ones(nep.n, dtype=complex)
    for _ in range(3):
        x = lu.solve(x)
        x /= np.linalg.norm(x)
    assert np.linalg.norm(M @ x) / abs(M).sum(axis=0).max() < 1e-13


def test_spmf_identities():
    # test/spmf.jl:37-63: compute_MM vs the closed form, and Mlincomb == from_MM == from_Mder
    rng = np.random.default_rng(0)
    n = 6
    A = [sp.random(n, n, 0.5, random_state=i, format="csc") for i in range(3)]
    nep = o.SPMF_NEP(A, [o.f_one, o.f_id, o.f_exp(-0.3)])
    S = rng.standard_normal((3, 3))
    V = rng.standard_normal((n, 3))
    import scipy.linalg as L
    Z = A[0] @ V + A[1] @ V @ S + A[2] @ V @ L.expm(-0.3 * S)
    assert np.linalg.norm(o.compute_MM(nep, S, V) - Z) < 1e-12
    a = np.array([1.0, 2.0, 0.5])
    lam = 0.3 + 0.1j
    z1 = o.compute_Mlincomb(nep, lam, V, a)
    z2 = o.compute_Mlincomb_from_MM(nep, lam, V, a)
    z3 = o.compute_Mlincomb_from_Mder(nep, lam, V, a)
    assert np.linalg.norm(z1 - z2) < 1e-12 and np.linalg.norm(z1 - z3) < 1e-12


def test_pep_dep_equal_spmf():
    # test/spmf.jl:158-178, :266-283
    rng = np.random.default_rng(1)
    n = 5
    A = [rng.standard_normal((n, n)) for _ in range(3)]
    pep = o.PEP(A)
    spmf = o.SPMF_NEP(o.get_Av(pep), o.get_fv(pep))
    V = rng.standard_normal((n, 4))
    a = np.array([1.0, 0.0, 3.0, 0.1])
    lam = -0.4 + 0.2j
    assert np.linalg.norm(o.compute_Mlincomb(pep, lam, V, a) - o.compute_Mlincomb(spmf, lam, V, a)) < 1e-12
    dep = o.nep_gallery("dep0")
    spmf = o.SPMF_NEP(o.get_Av(dep), o.get_fv(dep))
    V = rng.standard_normal((5, 4))
    assert np.linalg.norm(o.compute_Mlincomb(dep, lam, V, a) - o.compute_Mlincomb(spmf, lam, V, a)) < 1e-12
    assert np.linalg.norm(o.compute_Mder(dep, lam, 2) - o.compute_Mder(spmf, lam, 2)) < 1e-12


def test_startder_semantics():
    # test/core.jl:16-32
    dep = o.nep_gallery("dep0")
    rng = np.random.default_rng(2)
    V = rng.standard_normal((5, 3))
    lam = 0.7
    z = o.compute_Mlincomb(dep, lam, V, np.ones(3), startder=1)
    ref = sum(o.compute_Mder(dep, lam, j + 1) @ V[:, j] for j in range(3))
    assert np.linalg.norm(z - ref) < 1e-12


def test_structured_matrix_functions_match_generic_ones():
    # the oracle's closed form for f(lam*I + subdiag) against sqrtm / expm at a size where both are accurate
    import scipy.linalg as L
    k = 9
    lam = 3.0 + 0.5j
    S = np.diag(np.full(k, lam)) + np.diag(0.3 * np.arange(1, k), -1)
    F = o.f_isqrt_shift(0.7)(S)
    assert np.linalg.norm(F - 1j * L.sqrtm(S - 0.7 * np.eye(k))) < 1e-12 * np.linalg.norm(F)
    E = o.f_exp(-0.4)(S)
    assert np.linalg.norm(E - L.expm(-0.4 * S)) < 1e-12 * np.linalg.norm(E)


def test_oracle_resinv_dep0_literal():
    # docs/src/methods.md:18-19: resinv on dep0 converges to -0.15955391823299256
    from oracle import solvers as s
    nep = o.nep_gallery("dep0")
    lam, v = s.resinv(nep, lam=-0.2, v=np.ones(5), tol=1e-14)
    assert abs(lam - (-0.15955391823299256)) < 1e-13


def test_oracle_beyn_dep0_three_eigenvalues():
    # test/beyn.jl:34-37: exactly 3 eigenvalues in the disk of radius 1 centred at 0.2
    from oracle import solvers as s
    nep = o.nep_gallery("dep0")
    rng = np.random.default_rng(10)
    lam, V = s.contour_beyn(nep, rng.standard_normal((5, 5)), sigma=0.2, radius=1.0, N=1000, neigs=4, sanity_check=False)
    assert len(lam) == 3
    for l, v in zip(lam, V.T):
        assert np.linalg.svd(o.compute_Mder(nep, l), compute_uv=False)[-1] < 10000 * np.finfo(float).eps
        assert np.linalg.norm(o.compute_Mlincomb(nep, l, v)) / np.linalg.norm(v) < 10000 * np.finfo(float).eps


def test_oracle_iar_tiar_dep0():
    # test/iar.jl:23-37 (n=5, sigma... residual checks) and test/tiar.jl:59-70 (tiar == iar)
    from oracle import solvers as s
    nep = o.nep_gallery("dep0", 100)
    v0 = np.ones(100)
    lam, Q, V = s.iar(nep, sigma=0.0, gamma=1.0, neigs=3, maxit=60, v=v0, tol=1e-10)
    for l, q in zip(lam, Q.T):
        assert np.linalg.norm(o.compute_Mlincomb(nep, l, q)) / np.linalg.norm(q) < 1e-8
    G = V.conj().T @ V
    assert np.linalg.norm(G - np.eye(G.shape[0])) < 1e-6  # test/iar.jl:44-62
    lam2, Q2, Z, hist = s.tiar(nep, sigma=0.0, gamma=1.0, neigs=3, maxit=60, v=v0, tol=1e-10)
    assert np.allclose(np.sort_complex(lam), np.sort_complex(lam2), atol=1e-6)
    assert np.linalg.norm(Z.conj().T @ Z - np.eye(Z.shape[1])) < 1e-6


def test_oracle_block_SS_dep0():
    # test/contour_block_SS.jl:9-17
    from oracle import solvers as s
    nep = o.nep_gallery("dep0", 3)
    U, V = g.gen_rng_mat(g.MSWS_RNG(1), 3, 3), g.gen_rng_mat(g.MSWS_RNG(2), 3, 3)
    for radius in (1.0, (1.0, 2.0)):
        lam, Vec = s.contour_block_SS(nep, U, V, radius=radius, N=1000, sigma=0.1, K=3)
        assert np.linalg.norm(o.compute_Mlincomb(nep, lam[0], Vec[:, 0])) < np.sqrt(np.finfo(float).eps)


def test_oracle_iar_chebyshev_reference_cases():
    """test/iar_chebyshev.jl and the docstring of src/method_iar_chebyshev.jl:47-63."""
    from oracle import solvers as s
    eps = np.finfo(float).eps
    dep = o.nep_gallery("dep0")
    n = 5

    def check(nep, lam, Q, count, tol):
        assert len(lam) == count
        for l, q in zip(lam, Q.T):
            assert np.linalg.norm(o.compute_Mlincomb(nep, l, q)) / np.linalg.norm(q) < tol

    # "accuracy eigenpairs" (:87-90) and "Compute as many eigenpairs as possible" (:92-96): exactly 8 with maxit = 30
    lam, Q, err, V, H = s.iar_chebyshev(dep, sigma=0, neigs=5, maxit=100, tol=eps * 100, v=np.ones(n))
    check(dep, lam, Q, 5, n * np.sqrt(eps))
    assert np.linalg.norm(V.conj().T @ V - np.eye(V.shape[1]), 2) < n * np.sqrt(eps)  # "orthogonalization / DGKS" (:103-106)
    lam, Q, _, _, _ = s.iar_chebyshev(dep, sigma=0, neigs=np.inf, maxit=30, tol=eps * 100, v=np.ones(n))
    check(dep, lam, Q, 8, n * np.sqrt(eps))
    # the SPMF formula of compute_y0_cheb reproduces the DEP formula
    _, _, _, _, H1 = s.iar_chebyshev(dep, sigma=0, neigs=5, maxit=100, tol=eps * 100, v=np.ones(n))
    _, _, _, _, H2 = s.iar_chebyshev(dep, sigma=0, neigs=5, maxit=100, tol=eps * 100, v=np.ones(n), compute_y0_method="SPMF")
    assert H1.shape == H2.shape and np.abs(H1 - H2).max() < 1e-10
    # ... while it is stable: its DDf blocks grow like |D|^j / j!, and at k = 30 on [-1, 0] the SPMF formula has lost all
    # accuracy (same in the reference: the matrices are identical), whereas the DEP formula still resolves 8 eigenpairs
    lam_s, _, _, _, Hs = s.iar_chebyshev(dep, sigma=0, neigs=np.inf, maxit=30, tol=eps * 100, v=np.ones(n), compute_y0_method="SPMF")
    assert len(lam_s) < 8
    # docstring: dep0(100), tol = 1e-5, neigs = 3 prints these three eigenvalues
    dep100 = o.nep_gallery("dep0", 100)
    lam, Q, _, _, _ = s.iar_chebyshev(dep100, v=np.ones(100), tol=1e-5, neigs=3)
    for ref in (0.050462487848960284, -0.07708779190301127, 0.1503856540695659):
        assert np.min(np.abs(lam - ref)) < 1e-7
    # "Errors thrown" (:253-257)
    with pytest.raises(s.NoConvergenceException):
        s.iar_chebyshev(dep100, sigma=0, neigs=8, maxit=10, tol=eps * 100, v=np.ones(100))
    # "Scale Cheb's to different interval w DEP" (:78-83): neuron0, 10 eigenvalues on [-max tau, 0]
    neu = o.nep_gallery("neuron0")
    lam, Q, _, _, _ = s.iar_chebyshev(neu, a=-float(np.max(neu.tauv)), b=0, neigs=10, maxit=100, v=np.ones(2))
    check(neu, lam, Q, 10, 2 * np.sqrt(eps))
    # "DEP format with ComputeY0ChebSPMF_NEP" (:222-226): dep0_tridiag(1000), sigma = -1, gamma = 2
    tri = o.nep_gallery("dep0_tridiag", 1000)
    lam, Q, _, _, _ = s.iar_chebyshev(tri, sigma=-1, gamma=2, neigs=5, maxit=100, tol=eps * 100, compute_y0_method="SPMF", v=np.ones(1000))
    check(tri, lam, Q, 5, 1e-10)
    # "PEP" (:126-137): dense degree-3 PEP, n = 100 (matrices from the MSWS stream instead of rand)
    rng = g.MSWS_RNG(3)
    pep = o.PEP([(1 - g.gen_rng_mat(rng, 100, 100)) / 2 for _ in range(4)])
    lam, Q, _, _, _ = s.iar_chebyshev(pep, sigma=0, neigs=5, maxit=100, tol=eps * 100, v=np.ones(100))
    check(pep, lam, Q, 5, 100 * np.sqrt(eps))


def test_oracle_docstring_literals_dep0_100_and_qdep0():
    """More known answers of SURVEY.md 8(c): the tiar docstring eigenvalues of dep0(100) (src/method_tiar.jl:36-46; tol = 1e-5),
    the mslp eigenvalue of docs/src/index.md:64-67, and the qdep0 eigenvalue the quasinewton docstring converges to
    (src/errmeasure.jl:55-70).  (The three values printed in the iar docstring, src/method_iar.jl:37-40, belong to an older
    dep0 generator: iar and tiar are mathematically the same iteration and agree with each other here.)"""
    from oracle import solvers as s
    nep = o.nep_gallery("dep0", 100)
    v0 = np.ones(100)
    doc = np.array([0.050462487743188206, -0.07708769561361105, 0.1503916927814904])
    lam_t, _, _, _ = s.tiar(nep, v=v0, tol=1e-5, neigs=3)
    lam_i, _, _ = s.iar(nep, v=v0, tol=1e-5, neigs=3)
    for ref in doc:
        assert np.min(np.abs(lam_t - ref)) < 1e-8
        assert np.min(np.abs(lam_i - ref)) < 1e-8
    lam, v = s.resinv(nep, lam=0.0, v=v0, tol=1e-14)
    assert abs(lam - 0.05046248970129549) < 1e-12
    q = o.nep_gallery("qdep0")
    lam, v = s.resinv(q, lam=-1.0, v=np.ones(q.n), tol=1e-13, maxit=200)
    assert abs(lam - (-1.002466988585764)) < 1e-10


def test_c_restatement_of_compute_MM_matches_the_numpy_oracle():
    """oracle/csrc/spmf_mm.c (the all-core CPU baseline of bench.py): the reference's compute_MM loop (NEPTypes.jl:296-316)
    in C, checked against oracle.nep.compute_MM on gun and against a direct SciPy product on the synthetic stencil PEP."""
    import os
    import subprocess
    import scipy.sparse as sp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "oracle")], check=True, capture_output=True)
    from oracle.cspmf import CSpmf
    K, M, W1, W2 = g.load_gun_matrices()
    onep = o.nep_gallery("nlevp_native_gun")
    lam = 250.0 ** 2 + 1j
    f = [1.0, lam, 1j * np.sqrt(lam), 1j * np.sqrt(lam - 108.8774 ** 2)]
    rng = np.random.default_rng(0)
    V = rng.standard_normal((onep.n, 5)) + 1j * rng.standard_normal((onep.n, 5))
    Zo = o.compute_MM(onep, lam * np.eye(5), V)
    c = CSpmf([K, -M, W1, W2])
    for fn in (c.mm_csc, c.mm_csr):
        for threads in (1, 3):
            assert np.linalg.norm(fn(f, V, threads) - Zo) <= 1e-14 * np.linalg.norm(Zo)
    mats, _ = g.stencil_pep(40)
    lam = 0.3 + 0.2j
    c2 = CSpmf(mats)
    v = rng.standard_normal(1600) + 1j * rng.standard_normal(1600)
    Zr = sum(m * lam ** i for i, m in enumerate(mats)) @ v
    assert np.linalg.norm(c2.mm_csr([lam ** i for i in range(4)], v, 2)[:, 0] - Zr) <= 1e-14 * np.linalg.norm(Zr)
