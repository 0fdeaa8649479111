"""Host logic of the product's low-rank nleigs branch (nleigs.py: nleigs_lowrank / lowrank_backslash, the mirror of
src/method_nleigs.jl:380-518 with `P.is_low_rank` and src/rk_helper/rk_nep.jl:43-152) on the CPU: the device operator and the
device solver cache are replaced by the NumPy stand-ins of tests/host_standin.py, so what runs here is exactly the host code
the GPU test runs (tests/test_nleigs_gpu.py::test_nleigs_gun_lowrank_device), against the oracle and its golden file."""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as sla

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

import nepb200  # noqa: E402
from nepb200 import PowShift, ONE, IDENTITY, rk_helper  # noqa: E402
from host_standin import HostOperator  # noqa: E402
from oracle import gallery as g  # noqa: E402
from oracle import nep as o  # noqa: E402
from oracle import nleigs as onl  # noqa: E402
import make_nleigs_golden as mg  # noqa: E402
from make_nleigs_lowrank_golden import gun_lowrank_nep  # noqa: E402


class HostCache:
    """Stand-in for DeviceLinSolverCache: `get(shift, add_to_cache)` -> object with lin_solve; `solvers` dict."""

    def __init__(self, op):
        self.op, self.solvers, self.created = op, {}, 0

    def get(self, shift, add_to_cache=True):
        key = complex(shift)
        s = self.solvers.get(key)
        if s is None:
            M = sum(complex(f(key)) * A for f, A in zip(self.op.fi, self.op.A))
            lu = sla.splu(sp.csc_matrix(M, dtype=np.complex128))
            self.created += 1

            class S:
                def lin_solve(self, b, tol=0):
                    return lu.solve(np.asarray(b, dtype=np.complex128))
            s = S()
            if add_to_cache:
                self.solvers[key] = s
        return s


def _gun_host():
    K, M, W1, W2 = g.load_gun_matrices()
    fv = [ONE, IDENTITY, PowShift(0.5, 0.0, 1j), PowShift(0.5, 108.8774 ** 2, 1j)]
    return HostOperator([K, -M, W1, W2], fv), (K, M, W1, W2)


def test_low_rank_factors_and_descriptor():
    """low_rank_lu_factors / compactlu (rk_nep.jl:66-95) of the product equal the oracle's; ranks 19 + 65 = 84 for gun."""
    K, M, W1, W2 = g.load_gun_matrices()
    for A in (W1, W2):
        L, U = rk_helper.low_rank_lu_factors(A)
        Lo, Uo = o.low_rank_lu_factors(A)
        assert abs(L - Lo).max() == 0 and abs(U - Uo).max() == 0
        assert abs(L @ U.T - A).max() < 1e-14
    nep = nepb200.LowRankFactorizedNEP([W1, W2], [PowShift(0.5, 0.0, 1j), PowShift(0.5, 108.8774 ** 2, 1j)])
    assert nep.r == 84
    P = nepb200.LowRankStructure(1, nep.L, nep.U)
    assert (P.p, P.q, P.r) == (1, 2, 84) and P.Lcat.shape == (9956, 84) and P.UUt.shape == (84, 9956)
    # LowRankFactorizedNEP(L, U, f) (low_rank_nep.jl:32-43): A_i = L_i U_i'
    nep2 = nepb200.LowRankFactorizedNEP.from_factors(nep.L, nep.U, nep.fi)
    assert abs(nep2.A[0] - W1).max() < 1e-14 and abs(nep2.A[1] - W2).max() < 1e-14


def test_lowrank_backslash_matches_oracle():
    """One call of the product's lowrank_backslash against the oracle's backslash_ref on the same continuation vector
    (p = 1, N = 6: blocks of n, then r entries)."""
    op, _ = _gun_host()
    onep, _ = gun_lowrank_nep()
    Po = onl.RKNEP(onep)
    nep2 = onep.nep2
    P = nepb200.LowRankStructure(1, nep2.L, nep2.U)
    n, r, N, k = op.n, P.r, 6, 4
    rng = np.random.default_rng(3)
    sigma = 250.0 ** 2 + 5e4 * np.exp(2j * np.pi * rng.random(N + 2))
    xi = -(10.0 ** rng.uniform(2, 6, N + 2)) + 108.8774 ** 2
    beta = 0.5 + rng.random(N + 2)
    sgdd = (rng.standard_normal((4, N + 2)) + 1j * rng.standard_normal((4, N + 2))) / (1.0 + np.arange(N + 2))[None, :]
    wc = rng.standard_normal(n + N * r) + 1j * rng.standard_normal(n + N * r)
    cache = HostCache(op)

    def solve(s, y):
        return cache.get(s).lin_solve(y)
    from nepb200.nleigs import lowrank_backslash
    w = lowrank_backslash(op, P, solve, wc, sigma, k, beta, N, xi, sgdd)
    wo = onl.backslash_ref(wc, Po, solve, False, sigma, k, None, beta, N, xi, sgdd)
    assert np.linalg.norm(w - wo) <= 1e-12 * np.linalg.norm(wo)


def test_nleigs_lowrank_host_logic_variant_r1():
    """test/nleigs/nleigs_gun_variant_r1.jl through the product driver (stand-in operator and cache): 21 eigenvalues, equal to
    the oracle's golden run; leja = 0 with reusefact = 2 keeps one factorisation per distinct node (5)."""
    op, (K, M, W1, W2) = _gun_host()
    onep, _ = gun_lowrank_nep()
    P = nepb200.LowRankStructure(1, onep.nep2.L, onep.nep2.U)
    Sigma, Xi, nodes = mg.gun_setup()
    funres = mg.gun_residual(K, -M, W1, W2)
    v = mg.gun_start_vector(op.n)
    cache = HostCache(op)
    lam, X, res, det = nepb200.nleigs_lowrank(op, P, Sigma, Xi=Xi, maxit=100, v=v, leja=0, nodes=nodes, reusefact=2,
                                              errmeasure=funres, linsolvercache=cache)
    gold = json.load(open(os.path.join(HERE, "golden", "nleigs_gun_lowrank.json")))["R1"]
    assert len(lam) == 21 == gold["count"]
    gl = np.array([complex(*x) for x in gold["lam"]])
    for x in lam:
        assert np.min(np.abs(gl - x)) <= 1e-8 * abs(x)
    assert det["kconv"] == gold["kconv"] and det["N"] == gold["N"] and det["iterations"] == gold["iterations"]
    assert cache.created == 5 and np.all(res < 1e-10)
    assert det["rows"] == op.n + det["N"] * 84  # p = 1: one block of n, then blocks of r (method_nleigs.jl:205-211)
