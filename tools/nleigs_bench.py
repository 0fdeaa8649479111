"""Config C4 end to end: nleigs on the synthetic degree-3 stencil PEP (n = grid^2) with everything O(n) on the device.
Usage: nleigs_bench.py [grid=1000] [maxit=40] [nodes=2].  Prints one JSON line (wall-clock seconds; not a bench.py metric)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import nepb200
from nepb200 import synthetic, _lib

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 40
nnodes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
t0 = time.time()
mats, st = synthetic.stencil_pep(grid)
csc = [m.tocsc() for m in mats]
for m in csc:
    m.sort_indices()
dnep = nepb200.B200SPMF.from_nep(nepb200.PEP(csc))
t_build = time.time() - t0
t0 = time.time()
sym = nepb200.symbolic_info(dnep)
t_sym = time.time() - t0
centre, half = 0.96 - 0.6j, 0.06
Sigma = centre + half * np.array([-1 - 1j, -1 + 1j, 1 + 1j, 1 - 1j])
nodes = centre + half * 0.5 * np.exp(2j * np.pi * (np.arange(nnodes) + 0.25) / nnodes)  # fixed shifts after the linearization froze
v = np.ones(dnep.n) + 0j
l0 = _lib.lib.nepb_launch_count()
t0 = time.time()
lam, X, res, det = nepb200.nleigs(dnep, Sigma, v=v, maxit=maxit, nodes=nodes, minit=10)
_lib.lib.nepb_synchronize()
t_solve = time.time() - t0
print(json.dumps({"workload": "C4 stencil PEP deg 3, nleigs", "n": dnep.n, "nnz_union": dnep.nnz_union, "lu": sym, "build_s": round(t_build, 2),
                  "symbolic_s": round(t_sym, 2), "nleigs_s": round(t_solve, 2), "iterations": det["iterations"], "kconv": det["kconv"], "N": det["N"],
                  "factorizations_cached": det["factorizations"], "eigenvalues": [[float(x.real), float(x.imag)] for x in lam],
                  "residuals": [float(r) for r in res], "ritz_values_in_sigma": int(len(det["lam_all"])),
                  "best_residual_in_sigma": float(np.min(det["res_all"])) if len(det["res_all"]) else None, "gpu_launches": int(_lib.lib.nepb_launch_count() - l0)}))
