"""Dense tall-skinny blocks (DGKS orthogonalisation, DMMA products) and the device-resident iar / tiar loops against the
oracle.  Mirrors test/iar.jl:44-62 and test/tiar.jl:36-70 (orthogonality < 1e-6, tiar == iar) of the reference."""
import numpy as np
import pytest

import nepb200
from nepb200 import B200SPMF, Block, ONE, IDENTITY, PowShift
from oracle import gallery as g
from oracle import nep as o
from oracle import solvers as osol

pytestmark = pytest.mark.gpu


def gun_pair():
    K, M, W1, W2 = g.load_gun_matrices()
    dnep = B200SPMF([K, -M, W1, W2], [ONE, IDENTITY, PowShift(0.5, 0.0, 1j), PowShift(0.5, 108.8774 ** 2, 1j)])
    return o.nep_gallery("nlevp_native_gun"), dnep


@pytest.mark.parametrize("rows,k", [(1000, 1), (5000, 7), (40000, 33), (9956 * 6, 5), (300, 0)])
def test_dgks_matches_oracle(rows, k):
    rng = np.random.default_rng(rows + k)
    V, _ = np.linalg.qr(rng.standard_normal((rows, max(k, 1))) + 1j * rng.standard_normal((rows, max(k, 1))))
    V = V[:, :k]
    w = rng.standard_normal(rows) + 1j * rng.standard_normal(rows)
    if k:
        w += 50.0 * V @ (rng.standard_normal(k) + 0j)  # strong component in span(V): forces a second sweep
    Vb = Block(rows, k + 2)
    if k:
        Vb.upload(V, 0)
    Vb.upload(w, k)
    h, nrm, sweeps = nepb200.dgks(Vb, k, Vb, k)
    wo, ho = w.copy(), np.zeros(k, dtype=complex)
    nrmo = osol.orthogonalize_and_normalize_dgks(V, wo, ho)
    assert abs(nrm - nrmo) <= 1e-12 * nrmo
    assert np.linalg.norm(h - ho) <= 1e-12 * max(np.linalg.norm(ho), 1.0)
    wd = Vb.download(k, 1)[:, 0]
    assert np.linalg.norm(wd - wo) < 1e-11
    if k:
        assert np.linalg.norm(V.conj().T @ wd) < 1e-12 and sweeps >= 1
    assert abs(np.linalg.norm(wd) - 1) < 1e-13


@pytest.mark.parametrize("rows,ka,q", [(100, 1, 1), (1000, 5, 3), (9956, 20, 20), (20000, 101, 64), (777, 37, 70)])
def test_block_gemm_dmma(rows, ka, q):
    rng = np.random.default_rng(ka * q)
    A = rng.standard_normal((rows, ka + 3)) + 1j * rng.standard_normal((rows, ka + 3))
    Cm = rng.standard_normal((ka, q)) + 1j * rng.standard_normal((ka, q))
    Ab, Yb = Block.from_host(A), Block(rows, q + 1)
    nepb200.block_gemm(Ab, 2, ka, Cm, Yb, 1)
    Y = Yb.download(1, q)
    ref = A[:, 2:2 + ka] @ Cm
    assert np.linalg.norm(Y - ref) <= 1e-13 * np.linalg.norm(ref) * np.sqrt(ka)
    nrm = nepb200.colnorms(Yb, 1, q)
    assert np.allclose(nrm, np.linalg.norm(ref, axis=0), rtol=1e-12)


def test_solve_and_mlincomb_on_blocks():
    onep, dnep = gun_pair()
    n = dnep.n
    rng = np.random.default_rng(3)
    k = 6
    Y = rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k))
    a = (300.0 ** 2 - 200.0 ** 2) ** np.arange(k)
    a[0] = 0
    lam = 250.0 ** 2
    yb, tb = Block.from_host(Y), Block(n, 2)
    nepb200.mlincomb_block(dnep, lam, yb, 0, k, a, tb, 1)
    z = tb.download(1, 1)[:, 0]
    zo = o.compute_Mlincomb_from_Mder(onep, lam, Y, a)
    assert np.linalg.norm(z - zo) < 1e-11 * np.linalg.norm(zo)
    lu = nepb200.B200LU(dnep, [lam])
    nepb200.solve_block(lu, tb, 1, 1, yb, 0, alpha=-1.0)
    x = yb.download(0, 1)[:, 0]
    xo = -osol.FactorizeLinSolver(onep, lam).lin_solve(zo)
    assert np.linalg.norm(x - xo) < 1e-9 * np.linalg.norm(xo)


def test_tiar_device_equals_host_tiar_and_iar():
    A0, A1, tauv = g.dep0_matrices(100)
    dnep = B200SPMF.from_nep(nepb200.DEP([A0, A1], tauv))
    onep = o.nep_gallery("dep0", 100)
    v0 = np.ones(100)
    lam, Q, Z, hist = nepb200.tiar_device(dnep, sigma=0.0, neigs=3, maxit=60, v=v0, tol=1e-10)
    lo, Qo, Zo, histo = osol.tiar(onep, sigma=0.0, neigs=3, maxit=60, v=v0, tol=1e-10)
    assert len(lam) == len(lo) == 3
    assert np.allclose(np.sort_complex(lam), np.sort_complex(lo), atol=1e-9)
    assert np.linalg.norm(Z.conj().T @ Z - np.eye(Z.shape[1])) < 1e-6
    for l, q in zip(lam, Q.T):
        assert np.linalg.norm(o.compute_Mlincomb(onep, l, q)) / np.linalg.norm(q) < 1e-8
    lam2, Q2, V = nepb200.iar_device(dnep, sigma=0.0, neigs=3, maxit=60, v=v0, tol=1e-10)
    assert np.allclose(np.sort_complex(lam), np.sort_complex(lam2), atol=1e-6)
    assert np.linalg.norm(V.conj().T @ V - np.eye(V.shape[1])) < 1e-6


def test_iar_device_gun_matches_oracle():
    # config C2 at reduced depth (m = 30), everything resident in HBM
    onep, dnep = gun_pair()
    n = dnep.n
    kw = dict(sigma=250.0 ** 2, gamma=300.0 ** 2 - 200.0 ** 2, neigs=np.inf, maxit=30, v=np.ones(n), tol=1e-10, check_error_every=30)
    lam, Q, V = nepb200.iar_device(dnep, **kw)
    lo, Qo, Vo = osol.iar(onep, **kw)
    assert len(lam) == len(lo) and len(lam) >= 1
    a, b = np.sort_complex(lam), np.sort_complex(lo)
    assert np.max(np.abs(a - b) / np.abs(b)) < 1e-8
    assert np.linalg.norm(V.conj().T @ V - np.eye(V.shape[1])) < 1e-6
    lam_t, Qt, Zt, _ = nepb200.tiar_device(dnep, **kw)
    assert len(lam_t) == len(lam)
    assert np.max(np.abs(np.sort_complex(lam_t) - a) / np.abs(a)) < 1e-6


def test_iar_tiar_gun_at_m100_match_the_golden_ritz_values():
    """BASELINE config C2 at its stated depth m = 100.  The survey's gamma = 300^2 - 200^2 is not representable there (the
    reference forms gamma.^(0:m), method_iar.jl:77: 5e4^100 overflows Float64), so the depth-100 run uses gamma = 1000, the
    largest power of ten with a finite gamma^100; the converged Ritz values are pinned to the CPU oracle's
    (tests/golden/iar_gun_m100.json, written by tests/golden/make_iar_golden.py) and the residuals are recomputed here."""
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "iar_gun_m100.json")))
    lg = np.array(gold["lam_re"]) + 1j * np.array(gold["lam_im"])
    K, M, W1, W2 = g.load_gun_matrices()
    onep = o.nep_gallery("nlevp_native_gun")
    dnep = nepb200.B200SPMF([K, -M, W1, W2], [nepb200.ONE, nepb200.IDENTITY, nepb200.PowShift(0.5, 0.0, 1j),
                                              nepb200.PowShift(0.5, 108.8774 ** 2, 1j)])
    kw = dict(sigma=250.0 ** 2, gamma=1000.0, neigs=np.inf, v=np.ones(dnep.n), tol=1e-10, maxit=100, check_error_every=100)
    for fn in (nepb200.iar_device, nepb200.tiar_device):
        out = fn(dnep, **kw)
        lam, Q = out[0], out[1]
        assert len(lam) >= 8  # the oracle converges 10; pairs within a factor of the tolerance may fall on either side
        for x, q in zip(lam, Q.T):  # every returned pair is an eigenpair (pairs at the edge of the tolerance differ between runs)
            assert np.linalg.norm(o.compute_Mlincomb(onep, x, q)) / np.linalg.norm(q) < 1e-4  # absolute; |M| ~ 1e5
        # the best-converged golden values must all be found
        for x in lg[:6]:
            assert np.min(np.abs(lam - x)) < 1e-6 * abs(x)
