"""Regenerates tests/golden/nleigs_gun.json from the CPU oracle (oracle/nleigs.py): the reference's gun nleigs tests
(test/nleigs/nleigs_gun_naive.jl, nleigs_gun_variant_{p,r2,s}.jl with the set-up of test/rk_helper/gun_test_utils.jl:6-60),
full-rank SPMF branch.  The start vector is the MSWS stream 1-2u (seed 0) so that every implementation can rebuild it.
Run from the repo root:  python tests/golden/make_nleigs_golden.py [naive P R2 S]   (about 2-3 min of CPU per variant)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gallery as g  # noqa: E402
from oracle import nep as o  # noqa: E402
from oracle import nleigs as nl  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "nleigs_gun.json")


def gun_setup():
    """gun_test_utils.jl:6-34 (Sigma: closed half disk; 5 repeated nodes; pole candidates on the branch cut)."""
    gam, mu, sigma2 = 300.0 ** 2 - 200.0 ** 2, 250.0 ** 2, 108.8774
    xmin, xmax = mu - gam, mu + gam
    npts = 1000
    half = xmin + (xmax - xmin) * (np.exp(1j * np.linspace(0, np.pi, round(np.pi / 2 * npts) + 2)) / 2 + 0.5)
    Sigma = np.concatenate([half, [xmin]])
    nodes = gam * np.array([2 / 3, (1 + 1j) / 3, 0, (-1 + 1j) / 3, -2 / 3]) + mu
    Xi = -10 ** np.linspace(-8, 8, 10000) + sigma2 ** 2
    return Sigma, Xi, nodes


def gun_start_vector(n):
    rng = g.MSWS_RNG(0)
    return np.array([1 - 2 * g.gen_rng_float(rng) for _ in range(n)]) + 0j


def gun_residual(K, Mneg, W1, W2):
    """gun_test_utils.jl:45-60; `M` there is get_Av(PEP([K,-M]))[2] = -M."""
    nK, nM, nW1, nW2 = 1.474544889815002e+05, 2.726114618171165e-02, 2.328612251920476e+00, 3.793375498194695e+00
    s2 = 108.8774

    def f(lam, x):
        den = nK + abs(lam) * nM + np.sqrt(abs(lam)) * nW1 + np.sqrt(abs(lam - s2 ** 2)) * nW2
        return np.linalg.norm(K @ x + lam * (Mneg @ x) + 1j * np.sqrt(lam) * (W1 @ x) + 1j * np.sqrt(lam - s2 ** 2) * (W2 @ x)) / den
    return f


def run(variant):
    nep = o.nep_gallery("nlevp_native_gun")
    K, M, W1, W2 = g.load_gun_matrices()
    Sigma, Xi, nodes = gun_setup()
    funres = gun_residual(K, -M, W1, W2)
    t0 = time.time()
    if variant == "naive":
        sq = np.array([-1 - 1j, -1 + 1j, 1 + 1j, 1 - 1j])
        lam, X, res, d = nl.nleigs(nep, 150.0 ** 2 + 200.0 * sq, v=np.ones(nep.n) + 0j)
    else:
        v = gun_start_vector(nep.n)
        if variant == "P":
            lam, X, res, d = nl.nleigs(nep, Sigma, maxit=100, v=v, leja=0, nodes=nodes, reusefact=2, errmeasure=funres)
        elif variant == "R2":
            lam, X, res, d = nl.nleigs(nep, Sigma, Xi=Xi, minit=60, maxit=100, v=v, nodes=nodes, errmeasure=funres)
        elif variant == "S":
            lam, X, res, d = nl.nleigs(nep, Sigma, Xi=Xi, minit=70, maxit=100, v=v, nodes=nodes, static=True, errmeasure=funres)
        else:
            raise SystemExit("unknown variant " + variant)
    order = np.lexsort((lam.imag, lam.real))
    return {"count": int(len(lam)), "lam": [[float(x.real), float(x.imag)] for x in lam[order]], "res": [float(r) for r in res[order]],
            "kconv": int(d["kconv"]), "iterations": int(d["iterations"]), "N": int(d["N"]), "factorizations": int(d["factorizations"]),
            "oracle_seconds": round(time.time() - t0, 1)}


if __name__ == "__main__":
    todo = sys.argv[1:] or ["naive", "P", "R2", "S"]
    gold = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            gold = json.load(f)
    for var in todo:
        gold[var] = run(var)
        print(var, gold[var]["count"], gold[var]["kconv"], gold[var]["iterations"], gold[var]["oracle_seconds"], flush=True)
        with open(OUT, "w") as f:
            json.dump(gold, f, indent=1)
