"""Golden Ritz values of iar on gun at the deepest depth BASELINE config C2 names (m = 100), from the CPU oracle.
C2 as written in SURVEY.md 8(d) (gamma = 300^2 - 200^2 = 5e4) is not representable: the reference forms alpha = gamma.^(0:m)
(method_iar.jl:77), and 5e4^100 overflows Float64 -- in the reference as much as here.  The largest power of ten that keeps
gamma^100 finite is gamma = 1000; this script runs the oracle there and stores the converged Ritz values.
    python tests/golden/make_iar_golden.py        (about two minutes on one core)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import nep as o, solvers as osol  # noqa: E402

onep = o.nep_gallery("nlevp_native_gun")
n = onep.n
t0 = time.time()
kw = dict(sigma=250.0 ** 2, gamma=1000.0, neigs=np.inf, v=np.ones(n), tol=1e-10, maxit=100, check_error_every=100)
lam, Q, V = osol.iar(onep, **kw)
res = [float(np.linalg.norm(o.compute_Mlincomb(onep, l, q)) / np.linalg.norm(q)) for l, q in zip(lam, Q.T)]
out = {"config": "gun, iar sigma=250^2 gamma=1000 m=100 v=ones tol=1e-10 (default error measure)", "seconds": time.time() - t0,
       "lam_re": [float(x.real) for x in lam], "lam_im": [float(x.imag) for x in lam], "residual_norms": res}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "iar_gun_m100.json"), "w") as f:
    json.dump(out, f, indent=1)
print(out)
