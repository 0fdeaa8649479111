"""Regenerates tests/golden/nleigs_gun_lowrank.json from the CPU oracle: the reference's gun nleigs tests that go through
`LowRankFactorizedNEP` (test/rk_helper/gun_test_utils.jl:36-43: SumNEP(PEP([K, M]), LowRankFactorizedNEP([c1, c2]))), i.e. the
low-rank branches of `backslash` / `constructD` / `get_rk_nep` (src/method_nleigs.jl:380-518, src/rk_helper/rk_nep.jl:43-152):
variants R1 (leja = 0, reusefact = 2), R2 (minit = 60) and S (static, minit = 70); the reference asserts 21 eigenvalues for each
(test/nleigs/nleigs_gun_variant_{r1,r2,s}.jl).  Start vector: MSWS stream 1-2u (seed 0).
    python tests/golden/make_nleigs_lowrank_golden.py [R1 R2 S]      (R2 ~100 s, R1 / S ~20 s each)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import gallery as g  # noqa: E402
from oracle import nep as o  # noqa: E402
from oracle import nleigs as nl  # noqa: E402
from make_nleigs_golden import gun_setup, gun_start_vector, gun_residual  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "nleigs_gun_lowrank.json")


def gun_lowrank_nep():
    K, M, W1, W2 = g.load_gun_matrices()
    full = o.nep_gallery("nlevp_native_gun")
    return o.SumNEP(o.PEP([K, -M]), o.LowRankFactorizedNEP([W1, W2], o.get_fv(full.nep2))), (K, M, W1, W2)


def run(variant):
    nep, (K, M, W1, W2) = gun_lowrank_nep()
    Sigma, Xi, nodes = gun_setup()
    funres = gun_residual(K, -M, W1, W2)
    v = gun_start_vector(nep.n)
    t0 = time.time()
    if variant == "R1":
        lam, X, res, d = nl.nleigs(nep, Sigma, Xi=Xi, maxit=100, v=v, leja=0, nodes=nodes, reusefact=2, errmeasure=funres)
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
    gold = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for var in sys.argv[1:] or ["R1", "R2", "S"]:
        gold[var] = run(var)
        print(var, gold[var]["count"], gold[var]["oracle_seconds"], flush=True)
    with open(OUT, "w") as f:
        json.dump(gold, f, indent=1)
