"""Config C5 of BASELINE.json end to end: tiar on the waveguide problem in its native format (JARLEBRING, nx = nz + 4) with the
basis in HBM: compute_Mlincomb (stencil + chirp-z), the Schur-complement solve on the device multifrontal LU, DGKS and the
tall-skinny ZGEMMs.  Usage: wep_tiar_bench.py [nz] [maxit] [neigs]   (nz = 3*5*7*k odd; 945 = the survey's size, n = 898 695)"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import nepb200
from nepb200 import _lib
nz = int(sys.argv[1]) if len(sys.argv) > 1 else 315
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 50
neigs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
out = {"nz": nz, "nx": nz + 4}
t0 = time.perf_counter()
nep = nepb200.nep_gallery_WEP(nx=nz + 4, nz=nz, benchmark_problem="JARLEBRING", neptype="WEP")
out["n"] = nep.n
out["gallery_s"] = time.perf_counter() - t0
sigma = -3 - 3.5j
t0 = time.perf_counter()
solver = nepb200.WEPLinSolverCreator(solver_type="factorized").create_linsolver(nep, sigma)
_lib.lib.nepb_synchronize()
out["schur_assembly_analysis_factorisation_s"] = time.perf_counter() - t0
out["schur_lu"] = nepb200.symbolic_info(solver.schur)


class Creator:
    def create_linsolver(self, nep_, lam):
        return solver


v0 = np.ones(nep.n) / np.sqrt(nep.n)
l0 = _lib.lib.nepb_launch_count()
t0 = time.perf_counter()
try:
    lam, Q, Z, hist = nepb200.tiar_device(nep, sigma=sigma, neigs=neigs, maxit=maxit, v=v0, tol=1e-8, linsolvercreator=Creator())
    out["converged"] = True
except nepb200.NoConvergenceException as e:
    lam, Q = e.args[0], e.args[1]
    out["converged"] = False
_lib.lib.nepb_synchronize()
out["tiar_s"] = time.perf_counter() - t0
out["gpu_launches"] = int(_lib.lib.nepb_launch_count() - l0)
out["eigenvalues"] = [[float(x.real), float(x.imag)] for x in lam]
out["residuals"] = [float(np.linalg.norm(nep.compute_Mlincomb(lam[i], Q[:, i])) / np.linalg.norm(Q[:, i])) for i in range(len(lam))]
print(json.dumps(out))
