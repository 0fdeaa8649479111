"""The callers of the hot path (contour_beyn, iar, tiar, resinv) running on the B200 operators, against the CPU oracle.
Mirrors test/beyn.jl, test/iar.jl, test/tiar.jl, test/gun_native.jl and the C1 plumbing config of BASELINE.json.
Tolerance: eigenvalues / residuals within 1e-10 relative of the oracle (north star); sort permutations identical."""
import numpy as np
import pytest
import scipy.sparse as sp

import nepb200
from nepb200 import B200SPMF, ONE, IDENTITY, PowShift
from oracle import gallery as g
from oracle import nep as o
from oracle import solvers as osol

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def dep0_pair(n=5):
    A0, A1, tauv = g.dep0_matrices(n)
    return o.nep_gallery("dep0", n), B200SPMF.from_nep(nepb200.DEP([A0, A1], tauv))


def gun_pair():
    K, M, W1, W2 = g.load_gun_matrices()
    dnep = B200SPMF([K, -M, W1, W2], [ONE, IDENTITY, PowShift(0.5, 0.0, 1j), PowShift(0.5, 108.8774 ** 2, 1j)])
    return o.nep_gallery("nlevp_native_gun"), dnep


def msws_probe(n, k, seed=0):
    rng = g.MSWS_RNG(seed)
    return g.gen_rng_mat(rng, n, k)


def test_resinv_dep0_plumbing():
    # BASELINE config C1: dep0 n=5, lambda0=-0.2, v=ones, tol=1e-14 -> -0.15955391823299256 (docs/src/methods.md:18-19)
    onep, dnep = dep0_pair()
    lam, v = nepb200.resinv(dnep, lam=-0.2, v=np.ones(5), tol=1e-14)
    assert abs(lam - (-0.15955391823299256)) < 1e-13
    lo, vo = osol.resinv(onep, lam=-0.2, v=np.ones(5), tol=1e-14)
    assert abs(lam - lo) < 1e-13
    assert np.linalg.norm(o.compute_Mlincomb(onep, lam, v)) / np.linalg.norm(v) < 1e-13


def test_beyn_dep0_shifted_disk():
    # test/beyn.jl:32-46
    onep, dnep = dep0_pair()
    Vh = msws_probe(5, 5)
    lam, V = nepb200.contour_beyn(dnep, Vh, sigma=0.2, radius=1.0, N=1000, neigs=4, sanity_check=False, batch=250)
    assert len(lam) == 3
    lo, Vo = osol.contour_beyn(onep, Vh, sigma=0.2, radius=1.0, N=1000, neigs=4, sanity_check=False)
    # same eigenvalues; the order within the conjugate pair is a tie in |sigma - lambda| and may flip with rounding
    assert abs(lam[0] - lo[0]) < 1e-12 and {0, 1} == {int(np.argmin(abs(lam[1:] - x))) for x in lo[1:]}
    assert max(min(abs(lam - x)) for x in lo) < 1e-12
    for l, v in zip(lam, V.T):
        assert np.linalg.svd(o.compute_Mder(onep, l), compute_uv=False)[-1] < EPS * 10000
        assert np.linalg.norm(o.compute_Mlincomb(onep, l, v)) / np.linalg.norm(v) < EPS * 10000


def test_beyn_dep0_disk_at_origin_with_sanity_check():
    # test/beyn.jl:13-30, docstring example method_beyncontour.jl:36-42
    onep, dnep = dep0_pair()
    Vh = msws_probe(5, 3, seed=3)
    lam, V = nepb200.contour_beyn(dnep, Vh, radius=1.1, N=500, neigs=3, k=3 + 1 if False else 3)
    lo, Vo = osol.contour_beyn(onep, Vh, radius=1.1, N=500, neigs=3, k=3)
    assert len(lam) == len(lo)
    assert np.allclose(lam, lo, atol=1e-11)


def test_beyn_gun_reference_eigenvalue_and_moment_parity():
    # config C3 geometry at reduced N / k: sigma=150^2, radius=500 encloses exactly the eigenvalue of test/gun_native.jl:9
    onep, dnep = gun_pair()
    n = dnep.n
    Vh = msws_probe(n, 4)
    N = 16
    lam, V, A0, A1, info = nepb200.contour_beyn(dnep, Vh, sigma=150.0 ** 2, radius=500.0, N=N, neigs=3, k=4,
                                               batch=8, return_moments=True)
    lo, Vo, A0o, A1o, infoo = osol.contour_beyn(onep, Vh, sigma=150.0 ** 2, radius=500.0, N=N, neigs=3, k=4, return_moments=True)
    assert np.linalg.norm(A0 - A0o) / np.linalg.norm(A0o) < 1e-10
    assert np.linalg.norm(A1 - A1o) / np.linalg.norm(A1o) < 1e-10
    assert info["p"] == infoo["p"] == 1
    assert len(lam) == 1 and abs(lam[0] - (22345.116783765 + 0.644998598j)) < 10 ** -3.5  # test/gun_native.jl:18-19
    assert abs(lam[0] - lo[0]) / abs(lo[0]) < 1e-10
    r = np.linalg.norm(o.compute_Mlincomb(onep, lam[0], V[:, 0])) / np.linalg.norm(V[:, 0])
    ro = np.linalg.norm(o.compute_Mlincomb(onep, lo[0], Vo[:, 0])) / np.linalg.norm(Vo[:, 0])
    assert r < max(10 * ro, 1e-9 * abs(sp.csc_matrix(o.compute_Mder(onep, lam[0]))).sum(axis=0).max())


def test_beyn_sharded_nodes_sum_to_the_full_integral():
    # the multi-GPU seam without a communicator: two half-jobs (rank 0 / 1 of world 2) add up to the full moments
    onep, dnep = gun_pair()
    Vh = msws_probe(dnep.n, 3)
    kw = dict(sigma=150.0 ** 2, radius=500.0, N=8, neigs=2, k=3, return_moments=True, sanity_check=False)
    _, _, A0, A1, _ = nepb200.contour_beyn(dnep, Vh, **kw)
    integ = nepb200.ContourIntegrator(dnep, 3, 2, 4)
    h = 2 * np.pi / 8
    t = h * np.arange(8)
    gpt = 500.0 * (-np.sin(t) + 1j * np.cos(t))
    gt = 500.0 * (np.cos(t) + 1j * np.sin(t))
    W = np.stack([gpt * h / (2j * np.pi), gpt * gt * h / (2j * np.pi)], axis=1)
    parts = [integ.integrate(gt[r::2] + 150.0 ** 2, W[r::2], Vh)[0] for r in (0, 1)]
    S = parts[0] + parts[1]
    assert np.linalg.norm(S[:, :, 0] - A0) / np.linalg.norm(A0) < 1e-13
    assert np.linalg.norm(S[:, :, 1] - A1) / np.linalg.norm(A1) < 1e-13


def test_iar_dep0_matches_oracle():
    # test/iar.jl:23-37 at n=100 (shared start vector instead of randn)
    onep, dnep = dep0_pair(100)
    v0 = np.ones(100)
    lam, Q, V = nepb200.iar(dnep, sigma=0.0, neigs=3, maxit=60, v=v0, tol=1e-10)
    lo, Qo, Vo = osol.iar(onep, sigma=0.0, neigs=3, maxit=60, v=v0, tol=1e-10, errmeasure=None)
    assert len(lam) == len(lo) == 3
    assert np.allclose(np.sort_complex(lam), np.sort_complex(lo), atol=1e-9)
    for l, q in zip(lam, Q.T):
        assert np.linalg.norm(o.compute_Mlincomb(onep, l, q)) / np.linalg.norm(q) < 1e-8
    assert np.linalg.norm(V.conj().T @ V - np.eye(V.shape[1])) < 1e-6


def test_tiar_equals_iar_dep0():
    # test/tiar.jl:59-70
    onep, dnep = dep0_pair(100)
    v0 = np.ones(100)
    lam, Q, Z, hist = nepb200.tiar(dnep, sigma=0.0, neigs=3, maxit=60, v=v0, tol=1e-10)
    lam2, Q2, V = nepb200.iar(dnep, sigma=0.0, neigs=3, maxit=60, v=v0, tol=1e-10)
    assert np.allclose(np.sort_complex(lam), np.sort_complex(lam2), atol=1e-6)
    assert np.linalg.norm(Z.conj().T @ Z - np.eye(Z.shape[1])) < 1e-6
    with pytest.raises(nepb200.LostOrthogonalityException):
        nepb200.tiar(dep0_pair(5)[1], maxit=30, v=np.ones(5))  # method_tiar.jl:82-85
    with pytest.raises(nepb200.NoConvergenceException):
        nepb200.iar(dnep, sigma=0.0, neigs=30, maxit=5, v=v0)  # test/iar.jl:65-70


def test_iar_gun_short_run_matches_oracle():
    # config C2 at reduced depth: sigma=250^2, gamma=300^2-200^2, v=ones
    onep, dnep = gun_pair()
    n = dnep.n
    sigma, gamma = 250.0 ** 2, 300.0 ** 2 - 200.0 ** 2
    kw = dict(sigma=sigma, gamma=gamma, neigs=np.inf, maxit=25, v=np.ones(n), tol=1e-10, check_error_every=25)
    lam, Q, V = nepb200.iar(dnep, **kw)
    lo, Qo, Vo = osol.iar(onep, **kw)
    assert len(lam) == len(lo) and len(lam) >= 1
    a, b = np.sort_complex(lam), np.sort_complex(lo)
    assert np.max(np.abs(a - b) / np.abs(b)) < 1e-8
    for l, q in zip(lam, Q.T):
        assert np.linalg.norm(o.compute_Mlincomb(onep, l, q)) / np.linalg.norm(q) < 1e-6 * abs(l)


def test_contour_block_SS_dep0():
    # test/contour_block_SS.jl:9-17 (circle and ellipse, n = 3, k = 3, K = 3)
    A0, A1, tauv = g.dep0_matrices(3)
    dnep = B200SPMF.from_nep(nepb200.DEP([A0, A1], tauv))
    onep = o.nep_gallery("dep0", 3)
    U, V = msws_probe(3, 3, seed=1), msws_probe(3, 3, seed=2)
    for radius in (1.0, (1.0, 2.0)):
        lam, Vec, Shat, mp = nepb200.contour_block_SS(dnep, U, V, radius=radius, N=1000, sigma=0.1, k=3, K=3, batch=250, return_moments=True)
        lo, Veco, Shato, mpo = osol.contour_block_SS(onep, U, V, radius=radius, N=1000, sigma=0.1, K=3, return_moments=True)
        assert np.linalg.norm(Shat - Shato) <= 1e-11 * np.linalg.norm(Shato)
        assert mp == mpo
        assert np.linalg.norm(o.compute_Mlincomb(onep, lam[0], Vec[:, 0])) < np.sqrt(np.finfo(float).eps)
        assert max(min(abs(lam - x)) for x in lo) < 1e-8 * max(1.0, np.max(abs(lo)))


def test_ilan_device_matches_oracle():
    """ilan (src/method_ilan.jl) with the device operator: compute_Mlincomb, the shifted solve and
    `Bmult!` (one fused multi-term product with the k+1 x k+1 blocks G .* FDH_t) run on the B200, the recurrences on the host.
    The docstring example (dep_symm_double(10): 3 eigenpairs, first eigenvalue 0.03409997385842267, residual < 1e-5 as in
    test/ilan.jl:25-30) and the early Lanczos factorisation against the oracle; Ritz extraction from H as well."""
    import scipy.sparse as sp
    from nepb200 import B200SPMF, Monomial, ONE, Exp
    A, B, tau = g.dep_symm_double_matrices(10)
    n = A.shape[0]
    onep = o.DEP([A, B], tau)
    dnep = B200SPMF([-sp.identity(n, format="csc"), A, B], [Monomial(1), ONE, Exp(-tau[1])])
    lam, W, *_ = nepb200.ilan(dnep, v=np.ones(n), tol=1e-5, neigs=3)
    assert len(lam) == 3
    assert np.min(np.abs(lam - 0.03409997385842267)) < 1e-9
    for l, w in zip(lam, W.T):
        assert np.linalg.norm(o.compute_Mlincomb(onep, l, w)) / np.linalg.norm(w) < 1e-5
    A, B, tau = g.dep_symm_double_matrices(8)
    n = A.shape[0]
    onep = o.DEP([A, B], tau)
    dnep = B200SPMF([-sp.identity(n, format="csc"), A, B], [Monomial(1), ONE, Exp(-tau[1])])
    for proj_solve in (True, False):
        kw = dict(sigma=0.0, gamma=1.0, neigs=3, maxit=20, tol=1e-6, check_error_every=20, v=np.ones(n), proj_solve=proj_solve)
        lo, Wo, _, Vo, Ho, omo, HHo = osol.ilan(onep, errmeasure=o.residual_errmeasure(onep), **kw)
        lam, W, _, V, H, om, HH = nepb200.ilan(dnep, errmeasure=nepb200.ResidualErrmeasure(dnep), **kw)
        kk = 6  # the recurrence is not re-orthogonalised: rounding differences grow tenfold per step (tests/test_ilan.py)
        assert np.linalg.norm(V[:, :kk] - Vo[:, :kk]) < 1e-8
        assert np.linalg.norm(H[:kk, :kk] - Ho[:kk, :kk]) < 1e-8 * np.linalg.norm(Ho[:kk, :kk])
        assert np.linalg.norm(om[:kk] - omo[:kk]) < 1e-8 * np.linalg.norm(omo[:kk])
        assert len(lam) == len(lo) == 3
        for x in lam:
            assert np.min(np.abs(lo - x)) < 1e-7
        for l, w in zip(lam, W.T):
            assert np.linalg.norm(o.compute_Mlincomb(onep, l, w)) / np.linalg.norm(w) < 1e-6
    with pytest.raises(nepb200.NoConvergenceException):
        nepb200.ilan(dnep, neigs=2, maxit=3, tol=np.finfo(float).eps * 100, check_error_every=np.inf, v=np.ones(n))


def test_contour_c3_at_the_stated_size():
    """BASELINE config C3 at its full size: gun, contour_beyn sigma = 150^2, radius = 500, N = 128 nodes, k = 20 probe columns.
    The CPU oracle needs ~2 s per node, so the full integral is checked through what does not depend on the size --
      * sharding: the 8 round-robin shares of the nodes (what each of 8 GPUs integrates) add up to the 128-node moments (1e-12),
      * a 6-node sample of exactly these nodes / weights / probe against SuperLU solves at k = 20 (1e-10),
      * extraction: exactly one eigenvalue inside, equal to the reference's 22345.116783765 + 0.644998598im (test/gun_native.jl:9),
        with relative residual below 1e-10 -- and batch size / graph replay do not change a bit of the moments."""
    import scipy.sparse.linalg as sla
    onep, dnep = gun_pair()
    n, N, k = dnep.n, 128, 20
    Vh = msws_probe(n, k)
    sigma, radius = 150.0 ** 2, 500.0
    lam, V, A0, A1, info = nepb200.contour_beyn(dnep, Vh, sigma=sigma, radius=radius, N=N, neigs=5, k=k, batch=128, return_moments=True)
    assert len(lam) == 1 and info["p"] == 1
    assert abs(lam[0] - (22345.116783765 + 0.644998598j)) < 1e-6
    Mo = sp.csc_matrix(o.compute_Mder(onep, lam[0]))
    assert np.linalg.norm(Mo @ V[:, 0]) / (abs(Mo).sum(axis=0).max() * np.linalg.norm(V[:, 0])) < 1e-10
    h = 2 * np.pi / N
    t = h * np.arange(N)
    g_ = radius * (np.cos(t) + 1j * np.sin(t))
    gp = radius * (-np.sin(t) + 1j * np.cos(t))
    W = np.stack([gp * h / (2j * np.pi), gp * g_ * h / (2j * np.pi)], axis=1)
    integ = nepb200.ContourIntegrator(dnep, k, 2, 16)
    S = np.zeros((n, k, 2), dtype=complex)
    for r in range(8):
        part, flags = integ.integrate(g_[r::8] + sigma, W[r::8], Vh)
        assert not np.any(flags)
        S += part
    assert np.linalg.norm(S[:, :, 0] - A0) / np.linalg.norm(A0) < 1e-12
    assert np.linalg.norm(S[:, :, 1] - A1) / np.linalg.norm(A1) < 1e-12
    # the same 128 nodes in one batch, twice: bitwise reproducible
    integ2 = nepb200.ContourIntegrator(dnep, k, 2, 128)
    Sa, _ = integ2.integrate(g_ + sigma, W, Vh)
    Sb, _ = integ2.integrate(g_ + sigma, W, Vh)
    assert np.array_equal(Sa, Sb)
    assert np.linalg.norm(Sa[:, :, 0] - A0) / np.linalg.norm(A0) < 1e-12
    idx = np.array([0, 17, 31, 64, 99, 127])
    Ssub, _ = integ.integrate(g_[idx] + sigma, W[idx], Vh)
    ref = np.zeros_like(Ssub)
    for i in idx:
        X = sla.splu(sp.csc_matrix(o.compute_Mder(onep, g_[i] + sigma), dtype=complex)).solve(Vh.astype(complex))
        ref[:, :, 0] += W[i, 0] * X
        ref[:, :, 1] += W[i, 1] * X
    assert np.linalg.norm(Ssub - ref) / np.linalg.norm(ref) < 1e-10
    integ.close()
    integ2.close()
